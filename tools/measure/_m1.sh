#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -x -k "mu or trace or footprint or sweep" 2>&1 | tail -2
timeout 300 python tools/measure/mu_time.py cfg4 2>&1 | grep sweeps
timeout 300 python tools/measure/mu_time.py cfg2 2>&1 | grep sweeps
