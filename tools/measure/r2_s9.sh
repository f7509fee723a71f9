#!/bin/bash
mkdir -p gpurun_out/r2s9; cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc --launch-skip 1 -c 1 -o gpurun_out/r2s9/prof_gram_tc2 python tools/measure/mu_only.py cfg4 > gpurun_out/r2s9/ncu_tc.log 2>&1; echo ncu tc $?
