#!/bin/bash
mkdir -p gpurun_out/r2s7; cd /root/repo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s7/launches_cfg2_mu.csv python tools/measure/mu_only.py cfg2 > gpurun_out/r2s7/ncu_cfg2.log 2>&1; echo ncu cfg2 $?
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2s7/launches_cfg2_mu.csv | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc --launch-skip 1 -c 1 -o gpurun_out/r2s7/prof_gram_tc python tools/measure/mu_only.py cfg4 > gpurun_out/r2s7/ncu_tc.log 2>&1; echo ncu tc $?
tail -3 gpurun_out/r2s7/ncu_tc.log
