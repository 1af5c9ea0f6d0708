#!/bin/bash
# Round-2 baseline on a 2-GPU box with the round-1 kernels: the whole -m gpu suite (2-GPU tests included),
# the input-halo filters over NCCL, the sharded local_laplacian parity at 16K width, and the N=2 bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt 2>&1
HALIDE_B200_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_n2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_n2_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_rows_check.py > gpurun_out/r02_dist_rows_check_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dist_rows_check_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py 16384 2048 > gpurun_out/r02_dist_check_n2_r01kernels.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dist_check_n2_r01kernels.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload local_laplacian_16k > gpurun_out/r02_bench16k_r01kernels.json 2> gpurun_out/r02_bench16k_r01kernels.err
tail -3 gpurun_out/r02_n2_pytest.log; tail -2 gpurun_out/r02_dist_rows_check_n2.log; tail -2 gpurun_out/r02_dist_check_n2_r01kernels.log; cut -c1-600 gpurun_out/r02_bench16k_r01kernels.json
