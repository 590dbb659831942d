#!/bin/bash
# build variants on the GPU box (same image, nvcc present) and compare throughput / single-instance device phases
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
run() {
  echo "##### variant: $1"
  touch digital-subband-video-2_b200/csrc/dsvcu_api.cu
  make -s -j4 all EXTRA="$1" 2>&1 | grep -E "error|Error" | head -3
  GOPN=12 timeout 200 python tools/scale_probe.py 1,32 2>&1 | grep -E "^threads" | cut -c1-120
  DSV_PROFILE=2 GOPN=12 timeout 200 python tools/scale_probe.py 1 2>&1 | grep -E "device\]" | head -1 | cut -c1-200
}
for v in "$@"; do run "$v"; done
