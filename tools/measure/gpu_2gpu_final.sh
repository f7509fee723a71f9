#!/bin/bash
# 2-GPU bench line of the final build (the driver's own launch line)
mkdir -p gpurun_out/final2; cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/final2/bench_2gpu.json 2> gpurun_out/final2/bench_2gpu.err; echo rc $?
python - <<'PY'
import json
d=json.load(open("gpurun_out/final2/bench_2gpu.json")); print(d["value"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
