#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_full.log
tail -4 gpurun_out/r02_pytest_full.log
for n in 4k 16k_quarter 16k; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_$n > gpurun_out/r02_bench${n}_h.json 2> gpurun_out/r02_bench${n}_h.err
done
python - <<'P'
import json
for n in ("4k","16k_quarter","16k"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02_bench{n}_h.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}_h.err").read()[-1500:])
P
timeout 900 python tools/bench_all.py --quick > gpurun_out/r02_bench_all.log 2> gpurun_out/r02_bench_all.err
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_all.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["pipeline"], d["workload"], round(d["us_per_call"],1), "us", round(d["hbm_frac_of_measured"],3), {k:round(v["us"],1) for k,v in d.get("kernels",{}).items()})
P
