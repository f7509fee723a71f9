#!/bin/bash
# 2-GPU validation of the multi-rank bench path (torchrun, NCCL) + the 2-process GPU tests
mkdir -p gpurun_out/r2_2gpu_b; cd /root/repo
nvidia-smi topo -m > gpurun_out/r2_2gpu_b/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_2gpu_b/bench.json 2> gpurun_out/r2_2gpu_b/bench.err; echo bench2 $?; tail -5 gpurun_out/r2_2gpu_b/bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_2gpu_b/bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f n_gpus %d"%(d["value"], d["ms_per_step"], d["n_gpus"]))
print("e2e", {k:d["e2e"][k] for k in ("value","h2d_gbps_per_gpu","h2d_link_ceiling_gbps_per_gpu","host")})
print("e2e_res", d["e2e_resident"]["value"])
print("detail", json.dumps(d["scaling_detail"])[:900]); print("strong", d["strong_scaling"])
for n,l in (d["configs"] or {}).items():
    print(n, {k:l.get(k) for k in ("value","ms_per_step","frames_per_gpu","n_gpus")}, "frac", l.get("roofline",{}).get("frac"), "mu", {k:l.get("trace_update",{}).get(k) for k in ("stats_ms","sweeps_ms")}, l.get("error"))
P
