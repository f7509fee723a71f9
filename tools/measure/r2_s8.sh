#!/bin/bash
mkdir -p gpurun_out/r2s8; cd /root/repo
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k trace_statistics 2>&1 | grep -E "^\[|passed|failed|Error|error" | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q -x -k "mu_" 2>&1 | tail -2
python - <<'P' 2>&1 | tee gpurun_out/r2s8/timing.log
import torch, time, sys
sys.argv=["x","cfg4"]
exec(open("tools/measure/mu_only.py").read().split("if len(sys.argv) > 2")[0])
for rep in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); e.mu_stats(ids, beta); torch.cuda.synchronize()
    print("cfg4 40 frames stats wall ms %.2f"%((time.perf_counter()-t0)*1e3), "path", e.mu_path())
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_tc|stats_reduce" --csv --log-file gpurun_out/r2s8/launches_cfg4_mu.csv python tools/measure/mu_only.py cfg4 > gpurun_out/r2s8/ncu_cfg4.log 2>&1; echo ncu $?
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2s8/launches_cfg4_mu.csv | tail -4
