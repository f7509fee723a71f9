#!/bin/bash
# default bench line of the final state
mkdir -p gpurun_out/final; cd /root/repo
timeout 400 python bench.py > gpurun_out/final/bench2.json 2> gpurun_out/final/bench2.err; echo bench $?
python - <<'PY'
import json
d=json.load(open("gpurun_out/final/bench2.json"))
print(d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["trace_update"]["value"], d["reference_batch"]["value"], d["deformed_beta"]["value"], d["cpu_baseline"]["value"], d["clocks"])
PY
