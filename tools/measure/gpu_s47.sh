#!/bin/bash
# one-off: cfg1 and cfg4 lines of the final build
mkdir -p gpurun_out/s47; cd /root/repo
timeout 120 python bench.py --config cfg1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s47/bench_cfg1.json 2> gpurun_out/s47/bench_cfg1.err
timeout 200 python bench.py --config cfg4 --frames 100 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s47/bench_cfg4.json 2> gpurun_out/s47/bench_cfg4.err
python - <<'PY'
import json
for f in ("bench_cfg1","bench_cfg4"):
    try:
        d=json.load(open("gpurun_out/s47/%s.json"%f))
        print(f, d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["trace_update"], d["reference_batch"]["value"], d["deformed_beta"].get("value"))
    except Exception as e: print(f, "ERR", e)
PY
