#!/bin/bash
# round-end check after adopting the merged even/odd main loop: GPU suite, default bench, cfg3 line, smoke
mkdir -p gpurun_out/final2; cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final2/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/final2/pytest.log
timeout 400 python bench.py > gpurun_out/final2/bench.json 2> gpurun_out/final2/bench.err; echo bench $?
timeout 200 python bench.py --config cfg3 --frames 300 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/final2/bench_cfg3.json 2> gpurun_out/final2/bench_cfg3.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final2/smoke.log 2>&1; tail -1 gpurun_out/final2/smoke.log
python - <<'PY'
import json
for f in ("bench","bench_cfg3"):
    d=json.load(open("gpurun_out/final2/%s.json"%f))
    print(f, d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], (d["e2e"] or {}).get("value"), d["trace_update"]["value"], d["reference_batch"]["value"], d["deformed_beta"]["value"], d["deformed_beta"]["kernel_ms_per_launch"], (d["cpu_baseline"] or {}).get("value"))
PY
