#!/bin/bash
# The N-GPU record (torchrun, NCCL):  gpurun --gpus 8 --timeout 1200 -- 'bash tools/measure/record_ngpu.sh 8'
n=${1:-2}; out=gpurun_out/record_${n}gpu; mkdir -p $out; cd /root/repo
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $n --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; echo bench$n $?; tail -5 $out/bench.err
