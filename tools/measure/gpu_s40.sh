#!/bin/bash
# one-off: 2-GPU bench line of the final state (weak scaling sanity)
mkdir -p gpurun_out/s40; cd /root/repo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s40/bench_2gpu.json 2> gpurun_out/s40/bench_2gpu.err; echo rc $?
tail -3 gpurun_out/s40/bench_2gpu.err; cut -c1-400 gpurun_out/s40/bench_2gpu.json
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q -k "two_rank" 2>&1 | tail -3
