#!/bin/bash
cd "$GRAFT_REPO_ROOT/tools/micro" || exit 1
for T in 1 8 16 32 64; do ./launch_chain $T 4000 0 1; done
for T in 1 16 32; do ./launch_chain $T 4000 5 1; done
for T in 1 16 32; do ./launch_chain $T 2000 20 4; done
for T in 16 32; do ./launch_chain $T 2000 5 1 148; done
for T in 16 32; do ./launch_chain $T 2000 5 1 592; done
for T in 32; do ./launch_chain $T 1000 5 2000; done
