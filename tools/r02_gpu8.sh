#!/bin/bash
# every pipeline through the C ABI (timings + per-kernel split), then one ncu --set full capture per non-LL pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/bench_all.py > gpurun_out/r02_bench_all.log 2> gpurun_out/r02_bench_all.err
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_all.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["pipeline"], d["workload"], round(d["us_per_call"],1), "us", round(d["hbm_frac_of_measured"],3), {k:round(v["us"],1) for k,v in d.get("kernels",{}).items()})
P
prof() { # name W H regex
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$4" -s ${5:-3} -c ${6:-3} -o gpurun_out/r02_prof_$1 -f python tools/prof_run.py $1 $2 $3 3 > gpurun_out/r02_ncu_$1.log 2>&1
tail -1 gpurun_out/r02_ncu_$1.log
}
prof bilateral_grid 7680 4320 'bg_' 3 3
prof camera_pipe 2560 1920 'camera_pipe_kernel' 1 1
prof blur 7680 4320 'blur3x3' 1 1
prof stencil_chain 1536 2560 'stencil_chain' 8 1
prof nl_means 3840 2160 'nl_means' 1 1
