#!/bin/bash
# end-of-round evidence on one B200, most important first; later steps are skipped when the time budget runs out
# (usage: bash tools/r02_final2.sh [seconds])
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIMIT=${1:-170}
T0=$SECONDS
left() { echo $(( LIMIT - (SECONDS - T0) )); }
note() { echo "== $1 (t=$((SECONDS - T0))s)"; }
note pytest
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -6 gpurun_out/r02f_pytest.log | cut -c1-300
note bench16k
timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
note variants
timeout 60 python tools/ab_variants.py 2>&1 | tee gpurun_out/r02f_ab_variants.log
note smoke
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02f_smoke.log
python - <<'P'
import json
for n in ("r02f_bench",):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1]
        print(n, "ms/step", round(d["ms_per_step"],4), "Mpx/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pipelined"), {k:round(v["ms_per_step"],4) for k,v in d.get("kernels",{}).items()}, d.get("roofline") and round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/{n}.err").read()[-800:])
P
if [ $(left) -gt 40 ]; then
  note ncu16k
  timeout 60 ncu --set full --clock-control none --import-source on -k regex:'ll_level1_kernel|ll_up2_kernel' -s 8 -c 8 -o gpurun_out/r02f_prof_ll16k -f python tools/prof_run.py local_laplacian 16384 16384 2 > gpurun_out/r02f_ncu_16k.log 2>&1
  tail -1 gpurun_out/r02f_ncu_16k.log
fi
if [ $(left) -gt 35 ]; then
  note bench4k
  timeout 60 python bench.py --steps 20 --warmup 5 --workload local_laplacian_4k > gpurun_out/r02f_bench_4k.json 2> gpurun_out/r02f_bench_4k.err
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r02f_bench_4k.json') if l.startswith('{')][-1]
print('4k ms/step', round(d['ms_per_step'],4), 'Mpx/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('pipelined',{}).get('value'))"
fi
if [ $(left) -gt 55 ]; then
  note bench_all
  BENCH_ALL_VARIANT=2 timeout 100 python tools/bench_all.py 2> gpurun_out/r02f_bench_all.err | tee gpurun_out/r02f_bench_all.jsonl | cut -c1-260
fi
if [ $(left) -gt 25 ]; then
  note launches
  timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02f_launches_bench.log 2>&1
  tail -1 gpurun_out/r02f_launches.csv | cut -c1-200
fi
note done
