#!/bin/bash
# one-off: adaptive main-loop variant (full GPU suite, cfg2 / cfg3 bench)
mkdir -p gpurun_out/s45; cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s45/pytest.log 2>&1; echo pytest exit $?; tail -4 gpurun_out/s45/pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s45/bench_cfg2.json 2> gpurun_out/s45/bench_cfg2.err
timeout 200 python bench.py --config cfg3 --frames 300 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s45/bench_cfg3.json 2> gpurun_out/s45/bench_cfg3.err
DNMF_DYN_TAIL=0 timeout 200 python bench.py --config cfg3 --frames 300 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s45/bench_cfg3_static.json 2> gpurun_out/s45/bench_cfg3_static.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s45/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["deformed_beta"]["kernel_ms_per_launch"], d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
