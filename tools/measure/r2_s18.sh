#!/bin/bash
mkdir -p gpurun_out/r2s18; cd /root/repo
timeout 900 python -m pytest tests/test_gpu_edge.py -m gpu -q -x 2>&1 | tail -3
for t in "" "2,2,0,0,2,1" "1,1,0,0,2,4" "1,1,0,0,2,2" "2,1,0,0,2,1"; do
  DNMF_TILING=$t timeout 600 python tools/measure/fit_only.py cfg4 2>&1 | grep -v Warning | sed "s/^/[$t] /" | tee -a gpurun_out/r2s18/fit.log | cut -c1-200
done
