#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
K=${1:-ll_level1_kernel}
W=${2:-3840}; H=${3:-2160}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -o gpurun_out/r02_prof_${K}_${W} -f python tools/prof_run.py local_laplacian $W $H 3 > gpurun_out/r02_ncu_${K}.log 2>&1
tail -2 gpurun_out/r02_ncu_${K}.log
