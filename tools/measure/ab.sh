#!/bin/bash
# A/B timing of library variants (tools/measure/build_variant.sh): VARIANTS="a b default" CFGS=cfg2,cfg3 bash tools/measure/ab.sh <outdir>
out=gpurun_out/${1:-ab}; mkdir -p $out; cd /root/repo
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset DNMF_B200_LIB; else export DNMF_B200_LIB=$PWD/variants/$v/libdnmf_b200.so; fi
  timeout 600 python tools/measure/fit_only.py ${CFGS:-cfg2,cfg3} ${FRAMES:-0} 2>&1 | grep -v Warning | sed "s/^/$v: /" | tee -a $out/fit.log | cut -c1-110
done
