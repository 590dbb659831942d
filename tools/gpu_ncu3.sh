#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
TAG=${1:-n}
# level-0 launches of a P picture: prepass = 6th k_me_prepass of the 2nd frame, etc.
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_prepass -s 5 -c 1 -f -o $O/me_pre_$TAG python tools/prof_run.py 1920 1080 3 enc > $O/ncu_pre_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_subpel -s 0 -c 1 -f -o $O/me_sub_$TAG python tools/prof_run.py 1920 1080 3 enc > $O/ncu_sub_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_level -s 5 -c 1 -f -o $O/me_l0_$TAG python tools/prof_run.py 1920 1080 3 enc > $O/ncu_me_$TAG.log 2>&1
