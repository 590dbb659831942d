#!/bin/bash
# filter schedule variants (bands per CTA) built beforehand into digital-subband-video-2_b200/variants/:
# decode throughput with 1 and 32 instances, encode with 32
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
for v in "" "$@"; do
  if [ -n "$v" ]; then export DSV2CUDA_LIB=$PWD/digital-subband-video-2_b200/variants/libdsv2cuda_$v.so; fi
  echo "##### variant: ${v:-product}"
  timeout 300 python tools/scale_probe_dec.py 1,32 2>&1 | grep -E "^threads|^encode" | cut -c1-140
done
