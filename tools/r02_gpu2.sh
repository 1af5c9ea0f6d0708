#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_laplacian_gpu.py tests/test_golden_gpu.py tests/test_selftest_gpu.py -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -12 gpurun_out/r02_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload local_laplacian_4k > gpurun_out/r02_bench4k_b.json 2> gpurun_out/r02_bench4k_b.err
python - <<'P'
import json
for n in ("4k",):
    try:
        d=json.load(open(f"gpurun_out/r02_bench{n}_b.json"))
        print(n, "ms/step", d["ms_per_step"], "Mpx/s", d["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}_b.err").read()[-1500:])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ll_level1_pq|ll_final2|ll_down_pq|ll_up2' -s 12 -c 6 -o gpurun_out/r02_prof_ll_a python tools/prof_run.py local_laplacian 3840 2160 3 > gpurun_out/r02_ncu_a.log 2>&1
tail -3 gpurun_out/r02_ncu_a.log
