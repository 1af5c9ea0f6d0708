#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_laplacian_gpu.py tests/test_golden_gpu.py tests/test_selftest_gpu.py tests/test_blur_gpu.py -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -4 gpurun_out/r02_pytest.log
for n in 4k 16k; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload local_laplacian_$n > gpurun_out/r02_bench${n}_e.json 2> gpurun_out/r02_bench${n}_e.err
done
python - <<'P'
import json
for n in ("4k","16k"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02_bench{n}_e.json") if l.startswith("{")][-1]
        print(n, "ms/step", d["ms_per_step"], "Mpx/s", d["value"], "smooth", d["extra"]["smooth_frame_Mpixels_per_s"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e); print(open(f"gpurun_out/r02_bench{n}_e.err").read()[-1500:])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ll_up2_kernel|ll_level1_kernel' -s 5 -c 5 -o gpurun_out/r02_prof_ll_f -f python tools/prof_run.py local_laplacian 3840 2160 3 > gpurun_out/r02_ncu_f.log 2>&1
tail -2 gpurun_out/r02_ncu_f.log
