#!/bin/bash
# round 2, session 1: run the tcgen05 SYRK experiment for the first time, sanity of the round-1 suite, cfg3/cfg4 baselines
mkdir -p gpurun_out/r2s1; cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
for kv in 2688 1024 256; do timeout 60 tools/experiments/syrk_tf32_umma $kv 1184 > gpurun_out/r2s1/syrk_$kv.log 2>&1; echo "syrk $kv exit $?"; cat gpurun_out/r2s1/syrk_$kv.log; done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s1/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/r2s1/pytest.log
timeout 300 python bench.py --config cfg4 --frames 100 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s1/bench_cfg4.json 2> gpurun_out/r2s1/bench_cfg4.err; echo cfg4 $?
timeout 300 python bench.py --config cfg3 --frames 200 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s1/bench_cfg3.json 2> gpurun_out/r2s1/bench_cfg3.err; echo cfg3 $?
cut -c1-400 gpurun_out/r2s1/bench_cfg4.json; echo; cut -c1-400 gpurun_out/r2s1/bench_cfg3.json
