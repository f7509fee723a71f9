#!/bin/bash
# one-off: full GPU suite + deformed-beta throughput after the restage change
mkdir -p gpurun_out/s42; cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s42/pytest.log 2>&1; echo pytest exit $?; tail -4 gpurun_out/s42/pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s42/bench_cfg2.json 2> gpurun_out/s42/bench_cfg2.err
timeout 200 python bench.py --config cfg3 --frames 300 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s42/bench_cfg3.json 2> gpurun_out/s42/bench_cfg3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s42/bench_*.json")):
    try:
        d=json.load(open(f)); mu=d.get("trace_update") or {}; print(f, d["value"], d["roofline"]["kernel_ms_per_launch"], d["deformed_beta"]["value"], d["deformed_beta"]["kernel_ms_per_launch"], mu.get("stats_ms"), mu.get("sweeps_ms"))
    except Exception as e: print(f, "ERR", e)
PY
