#!/bin/bash
# round-2 diagnostics: current ME kernels, solo (full capture) and under load (app-range)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
MET=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__cache_requests_type_instruction_lookup_miss.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum.pct_of_peak_sustained_elapsed,sm__cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_level -s 5 -c 1 -f -o $O/me_l0_r2a python tools/prof_run.py 1920 1080 3 enc > $O/ncu_me_r2a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_prepass -s 5 -c 1 -f -o $O/me_pre_r2a python tools/prof_run.py 1920 1080 3 enc > $O/ncu_pre_r2a.log 2>&1
for T in 1 8 32; do
timeout 900 ncu --replay-mode app-range --clock-control none --section WarpStateStats --section SchedulerStats --metrics $MET -f -o $O/range_T${T}_r2a python tools/range_probe.py $T 3 > $O/ncu_range_T${T}_r2a.log 2>&1
done
timeout 600 python bench.py > $O/bench_r2a.json 2> $O/bench_r2a.err
tail -3 $O/ncu_range_T32_r2a.log
cat $O/bench_r2a.json | cut -c1-600
