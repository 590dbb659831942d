#!/bin/bash
# round-2 evidence: ncu --set full capture of every kernel family (one launch each, 1080p P picture
# of the second frame unless noted), launch list, range captures with all instances running
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=/tmp/prof; mkdir -p $O gpurun_out/prof
TAG=${1:-r2}
cap() { # name regex skip
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -f -o $O/${TAG}_$1 python tools/prof_run.py 1920 1080 3 ${4:-enc} > $O/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
}
cap me_prepass_L0 k_me_prepass 5
cap me_subpel k_me_subpel 0
cap me_level_L0 k_me_level 5
cap sbt_fwd_L1 k_sbt_fwd 6
cap sbt_inv_L1 k_sbt_inv 11
cap predict k_predict 0
cap reconstruct k_reconstruct 0
cap quant_hf_L2 k_quant_hf 5
cap compact_scatter k_compact_scatter 1
cap filter_skew k_filter_skew 1
cap pyr_interior k_pyr_interior 2
cap pyr_borders k_pyr_borders 2
cap dequant_hf k_dequant_hf 20 dec
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv python tools/prof_run.py 1920 1080 6 both > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
cp $O/${TAG}_launches.csv gpurun_out/prof/
python tools/ncu_summary.py $TAG r2 $O gpurun_out/prof
cp $O/${TAG}_me_level_L0.ncu-rep $O/${TAG}_filter_skew.ncu-rep gpurun_out/prof/ 2>/dev/null
du -sh gpurun_out/prof
