#!/bin/bash
# round 2, session 6: where does the statistics time go (cfg2 two-stage path, cfg4 tensor-core panel)
mkdir -p gpurun_out/r2s6; cd /root/repo
python - <<'P' 2>&1 | tee gpurun_out/r2s6/timing.log
import torch, time, numpy as np
from dnmf_b200.engine import Engine
from dnmf_b200.simulate import generate_video
def run(name, sz, K, T, sigma, shape_std, chunk):
    dev = torch.device("cuda:0")
    vid, positions, _ = generate_video(K, T, sz, shape_std, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=100, device=dev, frame_major=True)
    e = Engine(sz, K, T, dev)
    e.set_footprints(positions[:, :, 0], torch.full((K,), sigma), 3.5)
    e.upload_frames(vid.clamp_(min=0))
    beta = torch.zeros(10, 3, T, device=dev); beta[1, 0] = beta[2, 1] = beta[3, 2] = 1.0
    ids = torch.arange(T, dtype=torch.int32, device=dev)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(0, T, chunk):
            e.mu_stats(ids[i:i + chunk], beta)
        torch.cuda.synchronize(); print(name, "rep", rep, "stats wall ms %.2f" % ((time.perf_counter() - t0) * 1e3), "path", e.mu_path())
run("cfg2", [256, 128, 21], 150, 1000, 3.0, 3.0, 250)
run("cfg4", [256, 128, 21], 1000, 100, 6.0, 18.0, 100)
P
cat > /tmp/mu_only.py <<'P'
import torch, sys
from dnmf_b200.engine import Engine
from dnmf_b200.simulate import generate_video
cfg = sys.argv[1]
sz, K, T, sigma, ss = ([256, 128, 21], 150, 250, 3.0, 3.0) if cfg == "cfg2" else ([256, 128, 21], 1000, 40, 6.0, 18.0)
dev = torch.device("cuda:0")
vid, positions, _ = generate_video(K, T, sz, ss, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=100, device=dev, frame_major=True)
e = Engine(sz, K, T, dev)
e.set_footprints(positions[:, :, 0], torch.full((K,), sigma), 3.5)
e.upload_frames(vid.clamp_(min=0))
beta = torch.zeros(10, 3, T, device=dev); beta[1, 0] = beta[2, 1] = beta[3, 2] = 1.0
ids = torch.arange(T, dtype=torch.int32, device=dev)
e.mu_stats(ids, beta); e.mu_stats(ids, beta)
torch.cuda.synchronize()
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s6/launches_cfg2_mu.csv python /tmp/mu_only.py cfg2 > gpurun_out/r2s6/ncu_cfg2.log 2>&1; echo ncu cfg2 $?
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2s6/launches_cfg2_mu.csv | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc --launch-skip 1 -c 1 -o gpurun_out/r2s6/prof_gram_tc python /tmp/mu_only.py cfg4 > gpurun_out/r2s6/ncu_tc.log 2>&1; echo ncu tc $?
