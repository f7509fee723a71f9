"""Timing helper: dnmf_mu_stats over resident synthetic frames with CUDA events (python tools/measure/mu_time.py cfg4 [path flags])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dnmf_b200.engine import Engine  # noqa: E402
from dnmf_b200.simulate import generate_video  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
sz, K, T, sigma, ss = ([256, 128, 21], 150, 250, 3.0, 3.0) if cfg == "cfg2" else ([256, 128, 21], 1000, 100, 6.0, 18.0)
dev = torch.device("cuda:0")
vid, positions, _ = generate_video(K, T, sz, ss, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]},
                                   seed=100, device=dev, frame_major=True)
e = Engine(sz, K, T, dev)
e.set_footprints(positions[:, :, 0], torch.full((K,), sigma), 3.5)
e.upload_frames(vid.clamp_(min=0))
ids = torch.arange(T, dtype=torch.int32, device=dev)
if len(sys.argv) > 2:
    e.mu_path(int(sys.argv[2]))
for state in ("identity", "deformed"):
    beta = torch.zeros(10, 3, T, device=dev)
    beta[1, 0] = beta[2, 1] = beta[3, 2] = 1.0
    if state == "deformed":
        g = torch.Generator(device="cpu").manual_seed(5)
        beta[0, 0] += (2.0 / 128) * torch.randn(T, generator=g).to(dev)
        beta[0, 1] += (2.0 / 64) * torch.randn(T, generator=g).to(dev)
        beta[2, 0] += 0.005 * torch.randn(T, generator=g).to(dev)
    for _ in range(2):
        e.mu_stats(ids, beta)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(torch.cuda.current_stream(dev))
    for _ in range(3):
        e.mu_stats(ids, beta)
    b.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize()
    print("%s %s: mu_stats %.3f ms per %d frames, path bits %d" % (cfg, state, a.elapsed_time(b) / 3, T, e.mu_path()))
C = torch.rand(K, T, device=dev) + 0.1
for flags, label in ((0, "automatic"), (4, "one launch per sweep")):
    e.mu_path(flags)
    c = C.clone()
    e.mu_sweeps(c, None, 50)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(torch.cuda.current_stream(dev))
    for _ in range(3):
        e.mu_sweeps(c, None, 50)
    b.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize()
    print("%s: 50 sweeps (%s) %.3f ms per %d frames" % (cfg, label, a.elapsed_time(b) / 3, T))
e.mu_path(0)
c = C.clone()
for _ in range(2):
    e.mu_stats(ids, beta)
    e.mu_sweeps(c, None, 50)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(torch.cuda.current_stream(dev))
for _ in range(3):
    e.mu_stats(ids, beta)
    e.mu_sweeps(c, None, 50)
b.record(torch.cuda.current_stream(dev))
torch.cuda.synchronize()
print("%s: trace update (statistics + compaction + 50 sweeps) %.3f ms per %d frames" % (cfg, a.elapsed_time(b) / 3, T))
