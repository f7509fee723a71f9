#!/bin/bash
# one-off measurement script (round 1, session 3): sweeps with 4-aligned rows, statistics register budgets, 2-GPU sanity
mkdir -p gpurun_out/s37; cd /root/repo
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -k "mu_ or update_footprints or demo_trajectory or shuffled" > gpurun_out/s37/pytest.log 2>&1; echo pytest exit $?; tail -5 gpurun_out/s37/pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s37/bench_cfg2.json 2> gpurun_out/s37/bench_cfg2.err
for v in mumb10 mumb14; do DNMF_B200_LIB=/root/repo/variants/lib_$v.so timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s37/bench_cfg2_$v.json 2> gpurun_out/s37/bench_cfg2_$v.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s37/bench_*.json")):
    try:
        d=json.load(open(f)); mu=d.get("trace_update") or {}; print(f, d["value"], d["roofline"]["frac"], mu.get("stats_ms"), mu.get("sweeps_ms"), d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
