#!/bin/bash
# one-off: ncu --set full of the final fit kernel with a different deformation per frame (evidence for the next round)
mkdir -p gpurun_out/s48; cd /root/repo
DNMF_PROFILE_RANGE=deformed timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fit_tile_kernel -c 1 -o gpurun_out/s48/prof_deformed_v13 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s48/ncu.log 2>&1; echo rc $?; ls -la gpurun_out/s48
