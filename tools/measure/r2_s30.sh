#!/bin/bash
mkdir -p gpurun_out/r2s30; cd /root/repo
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_edge.py -m gpu -q -x -k "statistics or mu_stats or dense" 2>&1 | tail -3
python tools/measure/ext_only.py cfg2 250
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s30/launches_ext.csv python tools/measure/ext_only.py cfg2 250 > gpurun_out/r2s30/ncu.log 2>&1; echo ext $?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_tc|stats_reduce" --csv --log-file gpurun_out/r2s30/launches_cfg4_mu.csv python tools/measure/mu_only.py cfg4 > gpurun_out/r2s30/ncu_cfg4.log 2>&1; echo mu $?
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2s30/launches_cfg4_mu.csv | tail -4
