#!/bin/bash
# affine main loops: tests, base vs new timing, ncu --set full (identity + deformed) of the new build
mkdir -p gpurun_out/r2s13; cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for v in base default; do
  if [ $v = default ]; then unset DNMF_B200_LIB; else export DNMF_B200_LIB=$PWD/variants/$v/libdnmf_b200.so; fi
  timeout 600 python tools/measure/fit_only.py cfg2,cfg3,cfg4 2>&1 | grep -v Warning | tee gpurun_out/r2s13/fit_$v.log
done
unset DNMF_B200_LIB
for st in identity deformed; do
  DNMF_PROFILE_RANGE=cfg2_$st timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fit_tile -c 1 \
    -o gpurun_out/r2s13/prof_fit_$st python tools/measure/fit_only.py cfg2 > gpurun_out/r2s13/ncu_$st.log 2>&1; echo ncu $st $?
  ncu -i gpurun_out/r2s13/prof_fit_$st.ncu-rep --page raw --csv > gpurun_out/r2s13/fit_${st}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2s13/prof_fit_$st.ncu-rep --page source --csv > gpurun_out/r2s13/fit_${st}_src.csv 2>/dev/null
done
ls -la gpurun_out/r2s13
