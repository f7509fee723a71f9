#!/bin/bash
mkdir -p gpurun_out/r2s12; cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/sweep.sh 1 2>&1 | tail -14
