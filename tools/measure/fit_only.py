"""Kernel-experiment helper: the fused fit pass alone (dnmf_loss_grad = fit_tile_kernel + reduce_partials_kernel) and a
full-batch motion step, identity beta and a deformation per frame, ms per launch by CUDA events.

    [DNMF_B200_LIB=variants/<name>/libdnmf_b200.so] python tools/measure/fit_only.py cfg2[,cfg3,cfg4] [frames] [reps]

Prints one line per (config, state); also the checksum of the gradient and the loss so that two builds of the library
can be compared for bit-equality from the logs.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def run(name, frames, reps):
    cfg = bench.CONFIGS[name]
    T = frames or cfg.get("T_single", cfg["T"])
    dev = torch.device("cuda:0")
    tiling = tuple(int(v) for v in os.environ["DNMF_TILING"].split(",")) if os.environ.get("DNMF_TILING") else None
    dn, vid = bench.build_model(cfg, T, dev, 1, tiling)
    eng = dn.fp.engine
    beta = dn.fp.beta.detach()
    ids = torch.arange(T, dtype=torch.int32, device=dev)
    tl = eng.tiling()
    for state in ("identity", "deformed"):
        keep = bench.deform(beta, dn.affine, T, dev) if state == "deformed" else None
        ms, _ = bench.time_fit_kernel(eng, dn, beta, ids, reps, tag=name + "_" + state)
        g, sse = eng.loss_grad(ids, beta, dn.C)
        # full-batch motion step (what `value` times): fit + reduction + Adam
        opt = torch.optim.Adam([dn.fp.beta], lr=1e-9)
        _, st = dn._adam_state(opt)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        b2 = beta.clone()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(2):
            eng.motion_step(ids, b2, st["exp_avg"], st["exp_avg_sq"], dn.C, 1e-9, (0.9, 0.999), 1e-8, i + 1, dn.affine,
                            frames=None, B_global=T, loss_out=loss)
        torch.cuda.synchronize()
        ev0.record()
        for i in range(reps):
            eng.motion_step(ids, b2, st["exp_avg"], st["exp_avg_sq"], dn.C, 1e-9, (0.9, 0.999), 1e-8, i + 3, dn.affine,
                            frames=None, B_global=T, loss_out=loss)
        ev1.record()
        torch.cuda.synchronize()
        step_ms = ev0.elapsed_time(ev1) / reps
        print("%s %-8s T=%d fit %.4f ms  step %.4f ms  (%.3f us/frame)  sse %.10e  |g| %.10e  g[0:4] %.8e loss %.10e tiling %dx%dx%d z%d cap %d"
              % (name, state, T, ms, step_ms, 1e3 * ms / T, float(sse.sum()), float(g.double().norm()),
                 float(g[:4].double().abs().sum()), float(loss), tl["warps_x"], tl["warps_y"], tl["subtiles_y"], tl["warps_z"], tl["cap"]),
              flush=True)
        if keep is not None:
            beta.copy_(keep)
    eng.close()


if __name__ == "__main__":
    names = (sys.argv[1] if len(sys.argv) > 1 else "cfg2").split(",")
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    print("lib", os.environ.get("DNMF_B200_LIB", "default"), flush=True)
    for n in names:
        run(n, frames, reps)
