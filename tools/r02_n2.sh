#!/bin/bash
# 2-GPU validation of the row-sharded local_laplacian (exchange-free design) and of the input-halo filters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/r02_n${N}_gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 1000 640 > gpurun_out/r02_dist_check_n${N}_small.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dist_check_n${N}_small.log
tail -4 gpurun_out/r02_dist_check_n${N}_small.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py 16384 $((16384 / N)) > gpurun_out/r02_dist_check_n${N}.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dist_check_n${N}.log
tail -4 gpurun_out/r02_dist_check_n${N}.log
if [ "$N" = "2" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/dist_rows_check.py > gpurun_out/r02_dist_rows_check_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dist_rows_check_n2.log
tail -3 gpurun_out/r02_dist_rows_check_n2.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench16k_n${N}.json 2> gpurun_out/r02_bench16k_n${N}.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench16k_n${N}.json"))
    print("N=${N}", "ms/step", d["ms_per_step"], "Mpx/s", d["value"], "e2e", d["e2e"]["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench16k_n${N}.err").read()[-2500:])
P
