"""Profiling helper: a few device-resident shared-parameter iterations (extension) at cfg2, 250 frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 250
dev = torch.device("cuda:0")
dn, vid = bench.build_model(cfg, T, dev, 1)
dn.enable_shared_learning(lr_pos=1e-4, lr_sigma=1e-5, lr_background=1e-5)
opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
ids = torch.arange(T, dtype=torch.int32)
dn.update_motion(bench.Loader([(None, ids)] * 2), opt, epochs=1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
dn.update_motion(bench.Loader([(None, ids)] * 5), opt, epochs=1)
ev1.record()
torch.cuda.synchronize()
print("ms per shared step (T=%d): %.3f" % (T, ev0.elapsed_time(ev1) / 5))
