#!/bin/bash
# default bench line, reference arm, launch list, ncu --set full of the fit kernel (identity + deformed)
mkdir -p gpurun_out/r2s20; cd /root/repo
( time timeout 1200 python bench.py > gpurun_out/r2s20/bench.json 2> gpurun_out/r2s20/bench.err ) 2>&1 | grep real; echo bench $?
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s20/bench_ref.json 2> gpurun_out/r2s20/bench_ref.err ) 2>&1 | grep real; echo ref $?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s20/launches.csv python bench.py --steps 2 --warmup 3 --frames 200 --no-cpu-baseline --no-e2e --no-legs > gpurun_out/r2s20/ncu_list.log 2>&1; echo list $?
for st in identity deformed; do
  DNMF_PROFILE_RANGE=cfg2_$st timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fit_tile -c 1 \
    -o gpurun_out/r2s20/prof_fit_$st python tools/measure/fit_only.py cfg2 > gpurun_out/r2s20/ncu_$st.log 2>&1; echo ncu $st $?
  ncu -i gpurun_out/r2s20/prof_fit_$st.ncu-rep --page raw --csv > gpurun_out/r2s20/fit_${st}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2s20/prof_fit_$st.ncu-rep --page source --csv > gpurun_out/r2s20/fit_${st}_src.csv 2>/dev/null
done
DNMF_PROFILE_RANGE=cfg4_identity timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fit_tile -c 1 \
    -o gpurun_out/r2s20/prof_fit_cfg4 python tools/measure/fit_only.py cfg4 > gpurun_out/r2s20/ncu_cfg4.log 2>&1; echo ncu cfg4 $?
ncu -i gpurun_out/r2s20/prof_fit_cfg4.ncu-rep --page raw --csv > gpurun_out/r2s20/fit_cfg4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s20/prof_fit_cfg4.ncu-rep --page source --csv > gpurun_out/r2s20/fit_cfg4_src.csv 2>/dev/null
ls -la gpurun_out/r2s20
