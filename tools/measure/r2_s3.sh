#!/bin/bash
# round 2, session 3: whole GPU suite after the two-stage statistics / tensor-core panel / id checks
mkdir -p gpurun_out/r2s3; cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r2s3/pytest.log 2>&1; echo pytest exit $?; tail -25 gpurun_out/r2s3/pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r2s3/fullsize.log 2>&1; echo fullsize exit $?; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2s3/fullsize.log | cut -c1-700 | head -60
