#!/bin/bash
# final kernel: saturating u16 conversion + 48-row tiles: parity, then A/B (mask 128 = 32-row tiles)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_laplacian_gpu.py tests/test_selftest_gpu.py -x -q > gpurun_out/r02_pytest_13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_13.log
tail -4 gpurun_out/r02_pytest_13.log
timeout 200 python tools/ab_masks.py 16384 16384 0 128 0 128 2>&1 | tee gpurun_out/r02_ab13_16k.log
timeout 200 python tools/ab_masks.py 3840 2160 0 128 0 128 2>&1 | tee gpurun_out/r02_ab13_4k.log
timeout 200 python tools/ab_masks.py 16384 2048 0 128 2>&1 | tee gpurun_out/r02_ab13_band.log
