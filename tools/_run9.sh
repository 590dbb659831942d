timeout 200 python -m pytest tests/test_decode.py tests/test_encops.py -m gpu -x -q 2>&1 | tail -1
for v in "8 4" "16 8"; do set -- $v
if [ "$1" != "8" ]; then touch digital-subband-video-2_b200/csrc/k_filter.cuh; make -s EXTRA="-DFILT_LPC=$1 -DFILT_WPC=$2" 2>&1 | grep -v "^$" | head -3; fi
echo "== FILT_LPC=$1 FILT_WPC=$2 (edges split across lanes)"
DSV_PROFILE=2 GOPN=24 timeout 200 python tools/scale_probe.py 1,32 > gpurun_out/probe_fv.log 2>&1
grep -a "^threads" gpurun_out/probe_fv.log
grep -a -o "reconstruct+filters [0-9.]*" gpurun_out/probe_fv.log | awk '{v=$NF; s+=v; n++; if(n==1) f=v} END{print "recon+filters solo", f, "loaded avg", (s-f)/(n-1)}'
timeout 100 python tools/scale_probe_dec.py 1,32 2>&1 | tail -2 | cut -c1-60
done
