#!/bin/bash
# one-off: throughput with a different deformation per frame (late-fit state)
mkdir -p gpurun_out/s41; cd /root/repo
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s41/bench_cfg2.json 2> gpurun_out/s41/bench_cfg2.err
timeout 200 python bench.py --config cfg3 --frames 300 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s41/bench_cfg3.json 2> gpurun_out/s41/bench_cfg3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s41/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["roofline"]["kernel_ms_per_launch"], d["deformed_beta"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/s41/bench_cfg2.err
