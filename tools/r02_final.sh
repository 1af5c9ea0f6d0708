#!/bin/bash
# end-of-round evidence on one B200: full GPU suite, smoke, the bench lines (north star, 4K, reference arm),
# the ncu launch list of the bench command and one ncu --set full capture of the 16K frame's big kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -3 gpurun_out/r02_pytest_final.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --workload local_laplacian_4k > gpurun_out/r02_bench_4k.json 2> gpurun_out/r02_bench_4k.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python - <<'P'
import json
for n in ("r02_bench","r02_bench_4k","r02_bench_reference_arm"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],4) for k,v in d.get("kernels",{}).items()}, d.get("roofline") and round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/{n}.err").read()[-800:])
P
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_launches_bench.log 2>&1
tail -1 gpurun_out/r02_launches.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ll_level1_kernel|ll_up2_kernel|ll_down_rows' -s 12 -c 12 -o gpurun_out/r02_prof_ll16k -f python tools/prof_run.py local_laplacian 16384 16384 2 > gpurun_out/r02_ncu_16k.log 2>&1
tail -2 gpurun_out/r02_ncu_16k.log
