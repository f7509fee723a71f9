#!/bin/bash
# AFFK kernel + flat restage + 16-value epilogue: tests and A/B timing
mkdir -p gpurun_out/r2s14; cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for v in base noflat default; do
  if [ $v = default ]; then unset DNMF_B200_LIB; else export DNMF_B200_LIB=$PWD/variants/$v/libdnmf_b200.so; fi
  timeout 600 python tools/measure/fit_only.py cfg2,cfg3,cfg4 2>&1 | grep -v Warning | tee gpurun_out/r2s14/fit_$v.log
done
