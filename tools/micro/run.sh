#!/bin/bash
cd "$GRAFT_REPO_ROOT/tools/micro" || exit 1
for pb in 0 1024 3800; do for T in 1 32; do timeout 60 ./launch_chain $T 3000 5 1 $pb 0; done; done
for T in 1 32; do timeout 60 ./launch_chain $T 3000 5 1 0 2; done
for T in 1 32; do timeout 60 ./launch_chain $T 2000 50 8 1024 4; done
