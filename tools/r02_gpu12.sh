#!/bin/bash
# pair-LUT level-1 kernel + quad blur kernel: parity, then timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blur_gpu.py tests/test_local_laplacian_gpu.py tests/test_selftest_gpu.py -x -q > gpurun_out/r02_pytest_12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_12.log
tail -4 gpurun_out/r02_pytest_12.log
timeout 300 python tools/bench_all.py --only blur,local_laplacian 2> gpurun_out/r02_bench_all_12.err | tee gpurun_out/r02_bench_all_12.log | cut -c1-420
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_12.json 2> gpurun_out/r02_bench_12.err
python - <<'P'
import json
for n in ("r02_bench_12",):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],4) for k,v in d.get("kernels",{}).items()}, d.get("roofline") and round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/{n}.err").read()[-800:])
P
