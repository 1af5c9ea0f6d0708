#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_full.log
tail -25 gpurun_out/r02_pytest_full.log
