#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
for T in ${1:-1 32}; do
  echo "=== threads $T, DSV_PROFILE=1 (host phases)"
  DSV_PROFILE=1 GOPN=12 timeout 300 python tools/scale_probe.py $T 2>&1 | grep -E "threads|profile\]" | head -3 | cut -c1-420
  echo "=== threads $T, DSV_PROFILE=2 (device phases)"
  DSV_PROFILE=2 GOPN=12 timeout 300 python tools/scale_probe.py $T 2>&1 | grep -E "threads|device\]" | head -3 | cut -c1-300
  echo "=== threads $T, no profile"
  GOPN=12 timeout 300 python tools/scale_probe.py $T 2>&1 | grep -E "threads" | head -3 | cut -c1-300
done
