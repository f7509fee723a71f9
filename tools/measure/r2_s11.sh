#!/bin/bash
mkdir -p gpurun_out/r2s11; cd /root/repo
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -q -x 2>&1 | tail -25
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_tc|stats_reduce" --csv --log-file gpurun_out/r2s11/launches_cfg4_mu.csv python tools/measure/mu_only.py cfg4 > gpurun_out/r2s11/ncu_cfg4.log 2>&1; echo ncu $?
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2s11/launches_cfg4_mu.csv | tail -4
