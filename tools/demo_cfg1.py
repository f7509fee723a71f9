#!/usr/bin/env python
"""The reference's demo.py:16-46 schedule (K=10, T=100, 50x50x2, batch 4, Adam lr 1e-5,
5 x (10 epochs of update_motion + update_footprints(gamma_c=0, iter_c=50))) through the drop-in API,
timed end to end with host frames (the reference needs ~100 s for this on 8 CPU cores)."""
import os
import sys
import time

import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnmf_b200 import DeformableNMF, SimulatedVideoDataset  # noqa: E402


def main():
    torch.manual_seed(0)
    K, T = 10, 100
    sz = torch.tensor([50, 50, 2])
    dataset = SimulatedVideoDataset(K=K, T=T, sz=sz, shape_std=3, density=.2, bg_snr=-120, motion="gp", traces="exp",
                                    motion_par={"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=0)
    batch_size = 4
    dataloader = DataLoader(dataset, batch_size=batch_size, shuffle=True, num_workers=0)
    testloader = DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=0)
    dnmf = DeformableNMF(sz, K, T, positions=dataset.positions[:, :, 0], verbose=False)
    optimizer = torch.optim.Adam([dnmf.fp.beta], lr=1e-5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        dnmf.update_motion(dataloader, optimizer, gamma=1, epochs=10)
        A_t, Y_i, Y = dnmf.update_footprints(testloader, batch_size, sz, gamma_c=0, iter_c=50)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    losses = dnmf.losses()
    print("demo schedule: %.2f s for %d Adam steps (%d frame-iterations) + 5 x 50 trace sweeps; loss %.5f -> %.5f; "
          "A_t %s" % (dt, len(losses), 4 * len(losses), losses[0], losses[-1], A_t.shape))


if __name__ == "__main__":
    main()
