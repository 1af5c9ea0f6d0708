#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for hh in 2048 4096; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561 tools/shard1_time.py 16384 $hh 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee -a gpurun_out/r02_shard1_time.log
done
