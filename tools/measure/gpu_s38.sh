#!/bin/bash
# one-off measurement script (round 1, session 3): all sweeps of a frame in one CTA
mkdir -p gpurun_out/s38; cd /root/repo
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -k "mu_ or update_footprints or demo_trajectory or shuffled" > gpurun_out/s38/pytest.log 2>&1; echo pytest exit $?; tail -15 gpurun_out/s38/pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s38/bench_cfg2.json 2> gpurun_out/s38/bench_cfg2.err
timeout 200 python bench.py --config cfg3 --frames 300 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s38/bench_cfg3.json 2> gpurun_out/s38/bench_cfg3.err
timeout 100 python tools/demo_cfg1.py > gpurun_out/s38/demo_cfg1.log 2>&1; tail -5 gpurun_out/s38/demo_cfg1.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s38/bench_*.json")):
    try:
        d=json.load(open(f)); mu=d.get("trace_update") or {}; print(f, d["value"], d["roofline"]["frac"], mu.get("stats_ms"), mu.get("sweeps_ms"), d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
