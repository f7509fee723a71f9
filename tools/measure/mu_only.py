"""Profiling helper: two dnmf_mu_stats calls over resident synthetic frames (python tools/measure/mu_only.py cfg2|cfg4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dnmf_b200.engine import Engine  # noqa: E402
from dnmf_b200.simulate import generate_video  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
sz, K, T, sigma, ss = ([256, 128, 21], 150, 250, 3.0, 3.0) if cfg == "cfg2" else ([256, 128, 21], 1000, 40, 6.0, 18.0)
dev = torch.device("cuda:0")
vid, positions, _ = generate_video(K, T, sz, ss, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]},
                                   seed=100, device=dev, frame_major=True)
e = Engine(sz, K, T, dev)
e.set_footprints(positions[:, :, 0], torch.full((K,), sigma), 3.5)
e.upload_frames(vid.clamp_(min=0))
beta = torch.zeros(10, 3, T, device=dev)
beta[1, 0] = beta[2, 1] = beta[3, 2] = 1.0
ids = torch.arange(T, dtype=torch.int32, device=dev)
if len(sys.argv) > 2:
    e.mu_path(int(sys.argv[2]))
e.mu_stats(ids, beta)
e.mu_stats(ids, beta)
torch.cuda.synchronize()
print("path", e.mu_path())
