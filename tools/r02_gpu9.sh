#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_laplacian_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -3 gpurun_out/r02_pytest.log
for n in 4k 16k_quarter 16k; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_$n > gpurun_out/r02_bench${n}_g.json 2> gpurun_out/r02_bench${n}_g.err
done
python - <<'P'
import json
for n in ("4k","16k_quarter","16k"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02_bench{n}_g.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}_g.err").read()[-1500:])
P
bash tools/r02_gpu8.sh
