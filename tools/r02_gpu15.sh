#!/bin/bash
# bilateral_grid slice with conflict-free lane mapping; blur prefetch-depth A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bilateral_grid_gpu.py tests/test_blur_gpu.py -x -q > gpurun_out/r02_pytest_15.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_15.log
tail -4 gpurun_out/r02_pytest_15.log
timeout 200 python tools/ab_blur.py 7680 4320 0 2 0 2 32 48 2>&1 | tee gpurun_out/r02_ab15_blur.log
timeout 300 python tools/bench_all.py --only bilateral_grid 2> gpurun_out/r02_bench_all_15.err | tee gpurun_out/r02_bench_all_15.log | cut -c1-600
