#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=/tmp/range; mkdir -p $O gpurun_out/prof
TAG=${1:-r}
MET=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum.pct_of_peak_sustained_elapsed,sm__cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ctas_launched.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__registers_allocated.avg.pct_of_peak_sustained_elapsed,sm__registers_allocated.max.pct_of_peak_sustained_elapsed,l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,smsp__warps_launched.sum,sm__threads_launched.sum,gr__cycles_active.avg.pct_of_peak_sustained_elapsed,gr__ctas_launched_queue_sync.sum,gr__ctas_launched_queue_async.sum,fe__cycles_active.avg.pct_of_peak_sustained_elapsed,pcie__read_bytes.sum,pcie__write_bytes.sum
for T in ${2:-32}; do
timeout 900 ncu --replay-mode app-range --clock-control none --section WarpStateStats --section SchedulerStats --section Occupancy --metrics $MET -f -o $O/range_T${T}_$TAG python tools/range_probe.py $T 3 > $O/ncu_range_T${T}_$TAG.log 2>&1
tail -2 $O/ncu_range_T${T}_$TAG.log
ncu -i $O/range_T${T}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof/range_T${T}_$TAG.csv 2>/dev/null
done
