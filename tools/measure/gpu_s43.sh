#!/bin/bash
# one-off: ncu --set full of fit_tile_kernel with a different deformation per frame
mkdir -p gpurun_out/s43; cd /root/repo
DNMF_PROFILE_RANGE=deformed timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fit_tile_kernel -c 1 -o gpurun_out/s43/prof_deformed python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s43/ncu.log 2>&1; echo rc $?; tail -3 gpurun_out/s43/ncu.log; ls -la gpurun_out/s43
