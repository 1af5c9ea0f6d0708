#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
show() { python - "$1" <<'P'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print(sys.argv[1].split('/')[-1], "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "host", round(d.get("host_enqueue_ms_per_step",0),3), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
except Exception as e: print(sys.argv[1], "failed", e)
P
}
for n in 16k 16k_quarter; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 20 --warmup 5 --workload local_laplacian_$n > gpurun_out/r02_n${N}_$n.json 2>/dev/null
show gpurun_out/r02_n${N}_$n.json
done
