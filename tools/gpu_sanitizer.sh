#!/bin/bash
# compute-sanitizer over every kernel of an encode + decode (CIF, 4 frames: I + P pictures, all
# ME levels, filters, quantiser, transforms, pyramid) -- memcheck, racecheck, synccheck
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/sanitizer; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --log-file $O/r2_${tool}_cif.log python tools/prof_run.py 352 288 4 both > $O/r2_${tool}_cif.out 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $O/r2_${tool}_cif.log | tail -2 | tr '\n' ' ')"
done
# a second geometry with partial edge blocks and odd chroma
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --log-file $O/r2_memcheck_354x290.log python tools/prof_run.py 354 290 3 both > $O/r2_memcheck_354x290.out 2>&1
echo "memcheck 354x290 rc=$? : $(grep -E 'ERROR SUMMARY' $O/r2_memcheck_354x290.log | tail -1)"
