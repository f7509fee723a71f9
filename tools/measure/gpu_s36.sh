#!/bin/bash
# one-off measurement script (round 1, session 3): panel-kernel test, cfg4 4x4 vs 8x8 blocks, register-budget variants
mkdir -p gpurun_out/s36; cd /root/repo
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -k "mu_ or update_footprints or dense_neurons" > gpurun_out/s36/pytest.log 2>&1; echo pytest exit $?; tail -15 gpurun_out/s36/pytest.log
for m in 0 1; do DNMF_MU_BLOCK4=$m timeout 250 python bench.py --config cfg4 --frames 100 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s36/bench_cfg4_b$m.json 2> gpurun_out/s36/bench_cfg4_b$m.err; done
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s36/bench_cfg2.json 2> gpurun_out/s36/bench_cfg2.err
for v in 18 20; do DNMF_B200_LIB=/root/repo/variants/lib_minb$v.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s36/bench_cfg2_minb$v.json 2> gpurun_out/s36/bench_cfg2_minb$v.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s36/bench_*.json")):
    try:
        d=json.load(open(f)); mu=d.get("trace_update") or {}; print(f, d["value"], d["roofline"]["frac"], mu.get("stats_ms"), mu.get("sweeps_ms"), d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
