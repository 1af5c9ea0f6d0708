#!/bin/bash
# ncu --set full of the current level-1 / final kernels (8K band = quick) and of the blur and bilateral_grid kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ll_level1_kernel|ll_up2_kernel<\(bool\)1' -s 2 -c 2 -o gpurun_out/r02_prof_ll8k_c -f python tools/prof_run.py local_laplacian 8192 4096 2 > gpurun_out/r02_ncu_ll8k_c.log 2>&1
tail -2 gpurun_out/r02_ncu_ll8k_c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'blur3x3' -s 1 -c 1 -o gpurun_out/r02_prof_blur_c -f python tools/prof_run.py blur 7680 4320 2 > gpurun_out/r02_ncu_blur_c.log 2>&1
tail -2 gpurun_out/r02_ncu_blur_c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bg_' -s 3 -c 3 -o gpurun_out/r02_prof_bg_c -f python tools/prof_run.py bilateral_grid 7680 4320 2 > gpurun_out/r02_ncu_bg_c.log 2>&1
tail -2 gpurun_out/r02_ncu_bg_c.log
