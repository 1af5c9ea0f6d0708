#!/bin/bash
# why is the level-1 kernel ~30 % slower per row in multi-rank runs?  (a) one process alone on a 2-GPU box,
# (b) two independent single-GPU processes at once, (c) the 2-rank sharded run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print(sys.argv[1].split('/')[-1], "ms/step", round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, d["clocks"])
except Exception as e: print(sys.argv[1], "failed", e)
P
}
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_16k_quarter > gpurun_out/diag_a.json 2>/dev/null
show gpurun_out/diag_a.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_16k_quarter > gpurun_out/diag_b0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_16k_quarter > gpurun_out/diag_b1.json 2>/dev/null &
wait
show gpurun_out/diag_b0.json; show gpurun_out/diag_b1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --workload local_laplacian_16k_quarter > gpurun_out/diag_c.json 2>/dev/null
show gpurun_out/diag_c.json
nvidia-smi --query-gpu=index,power.limit,power.max_limit,clocks.max.sm,clocks.max.mem --format=csv
