#!/bin/bash
# single-GPU check of the round-2 local_laplacian kernels: parity suite, then the 4K and 16K bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -15 gpurun_out/r02_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload local_laplacian_4k > gpurun_out/r02_bench4k_a.json 2> gpurun_out/r02_bench4k_a.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_16k > gpurun_out/r02_bench16k_a.json 2> gpurun_out/r02_bench16k_a.err
python - <<'P'
import json
for n in ("4k","16k"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench{n}_a.json"))
        print(n, "ms/step", d["ms_per_step"], "Mpx/s", d["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}_a.err").read()[-1500:])
P
