#!/bin/bash
# measured parity numbers (tests print them) + ncu --set full of the shipped tensor-core panel
mkdir -p gpurun_out/r2s22; cd /root/repo
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_aux.py -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2s22/parity.txt; tail -3 gpurun_out/r2s22/parity.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc --launch-skip 1 -c 1 -o gpurun_out/r2s22/prof_gram_tc python tools/measure/mu_only.py cfg4 > gpurun_out/r2s22/ncu_tc.log 2>&1; echo ncu tc $?
ncu -i gpurun_out/r2s22/prof_gram_tc.ncu-rep --page raw --csv > gpurun_out/r2s22/tc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s22/prof_gram_tc.ncu-rep --page source --csv > gpurun_out/r2s22/tc_src.csv 2>/dev/null
