#!/bin/bash
# round 2, session 5: statistics timing after the two-stage rewrite (cfg2 fused tiles, cfg4 tensor-core vs SIMT panel)
mkdir -p gpurun_out/r2s5; cd /root/repo
timeout 300 python bench.py --config cfg4 --frames 100 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s5/bench_cfg4_tc.json 2> gpurun_out/r2s5/bench_cfg4_tc.err; echo cfg4 tc $?
DNMF_MU_PANEL=1 timeout 300 python bench.py --config cfg4 --frames 100 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s5/bench_cfg4_simt.json 2> gpurun_out/r2s5/bench_cfg4_simt.err; echo cfg4 simt $?
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s5/bench_cfg2.json 2> gpurun_out/r2s5/bench_cfg2.err; echo cfg2 $?
timeout 300 python bench.py --config cfg3 --frames 300 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s5/bench_cfg3.json 2> gpurun_out/r2s5/bench_cfg3.err; echo cfg3 $?
python - <<'P'
import json
for n in ("cfg4_tc","cfg4_simt","cfg2","cfg3"):
    try:
        d=json.loads(open("gpurun_out/r2s5/bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.4g frac %.3f"%(d["value"], d["roofline"]["frac"]), "mu", d["trace_update"], "deformed", d["deformed_beta"].get("value"))
    except Exception as e:
        print(n, "ERR", e)
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_tc|stats_reduce|mu_stats|fit_tile" -c 60 --csv --log-file gpurun_out/r2s5/launches_cfg4.csv python bench.py --config cfg4 --frames 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2s5/ncu_cfg4.log 2>&1; echo ncu $?
grep -E "gram_tc|stats_reduce" gpurun_out/r2s5/launches_cfg4.csv | awk -F'","' '{print $5, $NF}' | head -12
