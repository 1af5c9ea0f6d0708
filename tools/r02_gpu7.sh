#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_laplacian_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -3 gpurun_out/r02_pytest.log
for b in 1 2 4; do
HALIDE_B200_LL_COOP_BLOCKS_PER_SM=$b timeout 300 python bench.py --steps 20 --warmup 5 --workload local_laplacian_4k > gpurun_out/r02_bench4k_f$b.json 2> gpurun_out/r02_bench4k_f$b.err
done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench16k_f.json 2> gpurun_out/r02_bench16k_f.err
python - <<'P'
import json
for n in ("4k_f1","4k_f2","4k_f4","16k_f"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02_bench{n}.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}.err").read()[-1500:])
P
