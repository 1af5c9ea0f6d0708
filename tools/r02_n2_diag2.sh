#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print(sys.argv[1].split('/')[-1], "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
except Exception as e: print(sys.argv[1], "failed", e)
P
}
timeout 300 python -m pytest tests/test_local_laplacian_gpu.py tests/test_dist_gpu.py -x -q 2>&1 | tail -3
for n in 4k 16k; do
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_$n > gpurun_out/diag2_$n.json 2>/dev/null
show gpurun_out/diag2_$n.json
done
for n in 16k_quarter 16k; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --workload local_laplacian_$n > gpurun_out/diag2_n2_$n.json 2>/dev/null
show gpurun_out/diag2_n2_$n.json
done
