#!/usr/bin/env python
"""Benchmark of the dNMF fit hot path (BASELINE.json metric: frame-iterations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2]

One "step" = one Adam iteration of DeformableNMF.update_motion over a batch of B frames
(default B = all T frames of the rank's slab); one frame-iteration = forward +
loss + beta-gradient + Adam for one frame.  N > 1 is launched by torchrun (one rank per GPU); frames
are sharded (weak scaling: every rank owns T frames) and the only collective is an 8-byte loss
all-reduce per step.  The JSON line carries `value` (inputs resident in HBM), `e2e` (host buffers
through the public API, H2D inside the timed region), `roofline`, `cpu_baseline` and `clocks`.

`--impl reference` times the reference's CPU torch path (the oracle's TorchPort, which is
bit-identical to /root/reference's code -- tests/test_oracle.py) on the host cores, on the same
workload shape.  It is the only place besides cpu_baseline where bench.py executes oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (X, Y, Z, K, T per GPU, model sigma, simulator shape_std, deformation)
    "cfg1": dict(sz=(50, 50, 2), K=10, T=100, sigma=3.0, shape_std=3.0, deformation="quadratic",
                 desc="demo.py: 50x50x2, K=10, T=100, quadratic"),
    "cfg2": dict(sz=(256, 128, 21), K=150, T=1000, sigma=3.0, shape_std=3.0, deformation="affine",
                 desc="single-GPU synthetic volume 256x128x21, K=150, T=1000, affine"),
    "cfg3": dict(sz=(512, 256, 32), K=300, T=625, sigma=3.0, shape_std=3.0, deformation="quadratic", lr=1e-7,
                 desc="whole-brain 512x256x32, K=300, T=5000 over 8 GPUs (625 per GPU), quadratic"),
    "cfg4": dict(sz=(256, 128, 21), K=1000, T=250, sigma=6.0, shape_std=18.0, deformation="quadratic", lr=1e-7,
                 desc="dense stress 256x128x21, K=1000, sigma=6, T=2000 over 8 GPUs (250 per GPU)"),
}
CUTOFF = 3.5
LR = 1e-5   # demo.py:42; the raw-pixel quadratic basis needs a smaller step on large volumes (one step of 1e-5
            # moves a voxel at x=511 by 2.6 px through the x^2 term) -> per-config "lr"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the config's T)")
    ap.add_argument("--batch", type=int, default=0, help="frames per step and GPU (default: all T)")
    ap.add_argument("--tiling", default="", help="warps_x,warps_y,tz,cap override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per reference step (default 4; 1 for --impl reference)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference / CPU arm (oracle port of the reference's torch path)
# ------------------------------------------------------------------------------------------------
def make_cpu_workload(cfg, frames, seed=0):
    from dnmf_b200.simulate import generate_video
    vid, positions, _ = generate_video(cfg["K"], frames, cfg["sz"], cfg["shape_std"], .2, -120, "exp", "gp",
                                       {"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=seed, device="cpu",
                                       frame_major=True)
    return vid.clamp_(min=0).contiguous(), positions[:, :, 0].contiguous()


def time_reference(cfg, frames_per_step, steps, warmup):
    """Times the reference's CPU torch path (oracle.TorchPort) for `steps` Adam steps."""
    from oracle.dnmf_oracle import TorchPort
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    B = frames_per_step
    T = B * (steps + warmup)
    vid, pos0 = make_cpu_workload(cfg, T)
    port = TorchPort(cfg["sz"], cfg["K"], T, positions=pos0, shape_std=cfg["sigma"])
    opt = torch.optim.Adam([port.beta], lr=LR)
    affine = cfg["deformation"] == "affine"
    times = []
    for s in range(steps + warmup):
        ids = list(range(s * B, (s + 1) * B))
        t0 = time.perf_counter()
        port.motion_step(vid[ids], ids, opt, affine)
        times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    return B * len(timed) / sum(timed), sum(timed) / len(timed) * 1e3, cores


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_frames or 4
    value, ms, cores = time_reference(cfg, B, args.steps, args.warmup)
    sample = "%d frame(s) of the %s volume per Adam step through the reference's CPU torch path" % (B, args.config)
    line = {"impl": "reference", "metric": "frame-iterations/s", "value": value, "unit": "frame-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "frames_per_step": B, "lr": LR},
            "cpu_baseline": {"value": value, "unit": "frame-iterations/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "frame-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import torch.distributed as dist
    from dnmf_b200 import DeformableNMF, _lib
    from dnmf_b200.simulate import generate_video

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sz, K = cfg["sz"], cfg["K"]
    T = args.frames or cfg["T"]
    B = args.batch or T
    N = int(np.prod(sz))
    torch.manual_seed(1234 + rank)

    # synthetic slab of this rank, generated on the GPU (frame-major), then resident in the context
    vid, positions, _ = generate_video(K, T, sz, cfg["shape_std"], .2, -120, "exp", "gp",
                                       {"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=100 + rank, device=dev,
                                       frame_major=True)
    vid.clamp_(min=0)
    tiling = tuple(int(v) for v in args.tiling.split(",")) if args.tiling else None
    dn = DeformableNMF(sz, K, T, positions=positions[:, :, 0], cutoff=CUTOFF, deformation=cfg["deformation"],
                       shape_std=cfg["sigma"], device=dev, tiling=tiling, verbose=False, frame_offset=rank * T,
                       global_batch_scale=world)
    eng = dn.fp.engine
    dn.attach_video(vid, layout="TXYZ")
    host_frames = None
    if not args.no_e2e:
        host_frames = torch.empty((T,) + tuple(sz), dtype=torch.float32, pin_memory=True)
        host_frames.copy_(vid)
    del vid
    torch.cuda.empty_cache()

    opt = torch.optim.Adam([dn.fp.beta], lr=LR)
    _, st = dn._adam_state(opt)
    beta = dn.fp.beta.detach()
    ids_all = torch.arange(T, dtype=torch.int32, device=dev)
    batches = [ids_all[i:i + B] for i in range(0, T - B + 1, B)]
    # The loss of step i lands in slot i of a small ring.  Like the library (and the reference, which prints the
    # loss every 10th batch, Demix/dNMF.py:193), the ranks combine their partial losses every 10 steps: one
    # asynchronous all-reduce of a snapshot of the ring (the only collective of the reference-parity path), which
    # overlaps the following steps instead of synchronising the ranks.  Every outstanding reduction is waited
    # for inside the timed region.
    RING = 32
    REDUCE_EVERY = 10
    loss_ring = torch.zeros(RING, dtype=torch.float64, device=dev)
    works = []
    step_no = [0]
    last_slot = [0]
    reduced = [None]

    def reduce_losses():
        snap = loss_ring.clone()
        works.append((dist.all_reduce(snap, async_op=True), snap))

    def one_step(i, collective=True):
        ids = batches[i % len(batches)]
        step_no[0] += 1
        slot = step_no[0] % RING
        eng.motion_step(ids, beta, st["exp_avg"], st["exp_avg_sq"], dn.C, LR, (0.9, 0.999), 1e-8, step_no[0],
                        dn.affine, frames=None, B_global=B * world, loss_out=loss_ring[slot:slot + 1])
        if world > 1 and collective and step_no[0] % REDUCE_EVERY == 0:
            reduce_losses()
        last_slot[0] = slot

    def drain(final=False):
        if world > 1 and final:
            reduce_losses()                # the last steps' losses
        while works:
            w, snap = works.pop(0)
            w.wait()
            reduced[0] = snap

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    drain()
    barrier()
    c0 = eng.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the timed region must be long enough for nvidia-smi to sample clocks: repeat the K steps if short
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        one_step(args.warmup + i)
    drain(final=True)                      # the timed region ends after the last loss reduction
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    c1 = eng.counters()
    final_loss = float((reduced[0] if reduced[0] is not None else loss_ring)[last_slot[0]])
    # keep the GPU under the same load a little longer so the clock sampler sees it (not timed)
    t_end = time.time() + 1.5
    while rank == 0 and time.time() < t_end:
        one_step(0, collective=False)      # rank-local: the other ranks are not in this loop
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    st["step"].fill_(step_no[0])
    barrier()
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms)
    frame_iters = B * args.steps * world
    value = frame_iters / (ms_total * 1e-3)

    # ---- end to end: public API, host (pinned) frames, H2D + loss read-back inside the timed region ----
    e2e = None
    if host_frames is not None:
        Be = min(B, T)
        host_batches = [(host_frames[i:i + Be], torch.arange(i, i + Be, dtype=torch.int32))
                        for i in range(0, T - Be + 1, Be)]
        dn._video_resident = False

        class Loader:
            def __init__(self, items):
                self.items = items

            def __iter__(self):
                return iter(self.items)

            def __len__(self):
                return len(self.items)

        nb = len(host_batches)
        dn.update_motion(Loader([host_batches[i % nb] for i in range(max(1, min(args.warmup, 2)))]), opt, epochs=1)
        barrier()
        steps_e = args.steps
        loader = Loader([host_batches[i % nb] for i in range(steps_e)])
        ev0.record()
        dn.update_motion(loader, opt, epochs=1)          # each step ends with a D2H read of the loss
        ev1.record()
        barrier()
        t_e = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": Be * steps_e * world / (float(t_e) * 1e-3), "unit": "frame-iterations/s",
               "h2d_bytes_per_step": Be * N * 4 + Be * 4, "d2h_bytes_per_step": 8,
               "h2d_gbps_per_gpu": (Be * N * 4 + Be * 4) * steps_e / (float(t_e) * 1e-3) * 1e-9,
               "bound": "host-to-device link: every step ships its %.2f GB of fp32 frames from pinned host memory, "
                        "double-buffered against the fused kernel (compare h2d_gbps_per_gpu with the link rate)"
                        % (Be * N * 4 / 1e9),
               "api": "DeformableNMF.update_motion(host loader, torch.optim.Adam) -> dnmf_motion_step_host"}
        dn._video_resident = True

    if rank != 0:
        if world > 1:
            dist.barrier()                 # wait for rank 0's local roofline pass, then leave together
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fit_tile_kernel), timed live with CUDA events ----
    ids = batches[0]
    grad_scratch = torch.zeros(10, 3, T, dtype=torch.float32, device=dev)
    sse = torch.zeros(B, dtype=torch.float64, device=dev)
    lib = eng.lib
    import ctypes

    def fit_only():
        _lib.check(lib.dnmf_loss_grad(eng._h, None, ctypes.c_void_p(ids.data_ptr()), B, B,
                                      ctypes.c_void_p(beta.data_ptr()), ctypes.c_void_p(dn.C.data_ptr()),
                                      ctypes.c_void_p(grad_scratch.data_ptr()), ctypes.c_void_p(sse.data_ptr()),
                                      eng.stream), "dnmf_loss_grad")
    for _ in range(3):
        fit_only()
        torch.cuda.synchronize()       # lets the library's per-launch choice of the main-loop variant settle
    reps = max(5, args.steps)
    ev0.record()
    for _ in range(reps):
        fit_only()
    ev1.record()
    torch.cuda.synchronize()
    fit_ms = ev0.elapsed_time(ev1) / reps            # fused kernel + its (tiny) second-stage reduction

    rng = eng.ranges()
    ext = np.minimum(rng[:, :, 1] - rng[:, :, 0] + 2, np.asarray(sz)[None, :]).clip(min=0)
    k_eff = float(ext.prod(1).sum()) / N                       # true in-cutoff (voxel, neuron) pairs per voxel
    beta_id = torch.zeros(10, 3, T, device=dev)
    beta_id[1, 0], beta_id[2, 1], beta_id[3, 2] = 1.0, 1.0, 1.0
    counts, _, _, _ = eng.bin_tiles(beta_id, ids[:1])           # list lengths at the identity deformation
    tl = eng.tiling()
    listed = float((counts.astype(np.float64) * tl["tx"] * tl["ty"] * tl["tz"]).sum()) / N
    flops_per_frame = N * (144.0 + 25.0 * k_eff)
    bytes_per_frame = 4.0 * N
    peak_fp32 = ctypes.c_double(0.0)
    _lib.check(lib.dnmf_measure_fp32_peak(local, 5, ctypes.byref(peak_fp32)), "dnmf_measure_fp32_peak")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ach_tf = flops_per_frame * B / (fit_ms * 1e-3) * 1e-12
    ach_gbs = bytes_per_frame * B / (fit_ms * 1e-3) * 1e-9
    traffic, traffic_src = None, None
    try:   # DRAM bytes per frame of fit_tile_kernel from the committed `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "fit_tile_traffic.json")))
        if tj.get("config") == args.config:
            traffic, traffic_src = tj["dram_bytes_per_frame"] * B, tj["source"]
    except Exception:
        pass
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": peak_fp32.value, "unit": "TFLOP/s",
                "frac": ach_tf / peak_fp32.value if peak_fp32.value else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_frame * B,
                "kernel": "fit_tile_kernel", "kernel_ms_per_launch": fit_ms, "frames_per_launch": B,
                "flops_per_frame_iter": flops_per_frame, "k_eff_in_cutoff_pairs_per_voxel": k_eff,
                "listed_pairs_per_voxel": listed,
                "peak_source": "dnmf_measure_fp32_peak (FFMA microbenchmark, measured live in this run)",
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "bytes_per_frame_iter": bytes_per_frame,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}}

    # ---- secondary metric: the same fused pass late in a fit, when every frame has its own deformation ----
    # (at the start of a fit all beta_t are the identity, so consecutive frames of a CTA share the window, the
    # neuron list and the staged slices; here each frame gets its own translation / affine / quadratic terms and
    # the CTA rebuilds list and slices for every frame)
    deformed = None
    try:
        gen = torch.Generator().manual_seed(7)
        scale = torch.tensor([2.0, 5e-3, 5e-3, 5e-3, 1e-5, 1e-5, 1e-5, 1e-5, 1e-5, 1e-5])[:, None, None]
        if dn.affine:
            scale[4:] = 0.0
        scale = scale * torch.tensor([1.0, 1.0, 0.25])[None, :, None]   # z is shallow: a quarter of the x, y motion
        beta_keep = beta.clone()
        beta.add_((scale * torch.randn(10, 3, T, generator=gen)).to(dev))
        for _ in range(3):
            fit_only()
            torch.cuda.synchronize()
        if os.environ.get("DNMF_PROFILE_RANGE") == "deformed":   # ncu --profile-from-start off: capture this launch only
            torch.cuda.cudart().cudaProfilerStart()
            fit_only()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        ev0.record()
        for _ in range(reps):
            fit_only()
        ev1.record()
        torch.cuda.synchronize()
        d_ms = ev0.elapsed_time(ev1) / reps
        beta.copy_(beta_keep)
        deformed = {"value": B / (d_ms * 1e-3), "unit": "frame-iterations/s", "kernel_ms_per_launch": d_ms,
                    "what": "fit_tile_kernel + reduction with a different deformation per frame (translations "
                            "sigma = 2 px in x, y and 0.5 px in z, linear terms 0.5 %, no two frames alike): list and "
                            "slices rebuilt for every frame"}
    except Exception as ex:  # pragma: no cover
        deformed = {"error": repr(ex)}

    # ---- secondary metric: the reference's own minibatch size (demo.py: 4 frames) through the public API ----
    ref_batch = None
    if T >= 8:
        try:
            Bq = 4
            ids_cpu = torch.arange(T, dtype=torch.int32)
            loader4 = [(None, ids_cpu[i:i + Bq]) for i in range(0, T - Bq + 1, Bq)]
            dn.update_motion(loader4[:8], opt, epochs=1)       # warm-up
            torch.cuda.synchronize()
            ev0.record()
            dn.update_motion(loader4, opt, epochs=1)
            ev1.record()
            torch.cuda.synchronize()
            ref_batch = {"batch": Bq, "steps": len(loader4), "unit": "frame-iterations/s",
                         "value": len(loader4) * Bq / (ev0.elapsed_time(ev1) * 1e-3),
                         "api": "DeformableNMF.update_motion(4-frame minibatches over the attached video) -> "
                                "dnmf_motion_epoch"}
        except Exception as ex:  # pragma: no cover
            ref_batch = {"error": repr(ex)}

    # ---- secondary metric: trace update (update_footprints hot loop #2), frame-MU-iterations/s ----
    mu = None
    if not args.no_mu:
        try:
            iters = 50                                    # demo.py:46 uses iter_c=50
            eng.mu_stats(ids_all[:min(T, 8)], beta)       # warm-up: allocations, lazy module load
            eng.mu_sweeps(dn.C.clone(), 0.0, 1)
            torch.cuda.synchronize()
            ev0.record()
            for i in range(0, T, 250):
                eng.mu_stats(ids_all[i:i + 250], beta)
            ev1.record()
            torch.cuda.synchronize()
            stats_ms = ev0.elapsed_time(ev1)
            Cmu = dn.C.clone()
            ev0.record()
            eng.mu_sweeps(Cmu, 0.0, iters)
            ev1.record()
            torch.cuda.synchronize()
            sweeps_ms = ev0.elapsed_time(ev1)
            mu = {"metric": "frame-MU-iterations/s", "value": T * iters / ((stats_ms + sweeps_ms) * 1e-3),
                  "frames": T, "iter_c": iters, "stats_ms": stats_ms, "sweeps_ms": sweeps_ms,
                  "what": "dnmf_mu_stats over all frames + %d multiplicative sweeps (update_footprints without "
                          "the dense returns); reference: 1 035 frame-MU-iters/s on the 8-core CPU at cfg1" % iters}
        except Exception as exc:                          # keep the headline line even if this leg fails
            mu = {"error": str(exc)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        Bc = args.cpu_frames or 4
        v, ms, cores = time_reference(cfg, Bc, 2, 1)
        cpu_baseline = {"value": v, "unit": "frame-iterations/s", "cores": cores, "kind": "port",
                        "sample": "2 Adam steps of %d frames of the %s volume (after 1 warm-up) through the "
                                  "oracle's torch port of the reference CPU path" % (Bc, args.config),
                        "ms_per_step": ms}

    launches = sum(c1[k] - c0[k] for k in ("fit_launches", "reduce_launches", "adam_launches"))
    line = {"metric": "frame-iterations/s", "value": value, "unit": "frame-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "name": args.config, "frames_per_gpu": T, "frames_per_step_per_gpu": B,
                       "cutoff_sigma": CUTOFF, "lr": LR, "tiling": tl,
                       "l2": "inputs (%.2f GB per step) larger than L2" % (B * N * 4 / 1e9)},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "final_loss": final_loss, "trace_update": mu,
            "reference_batch": ref_batch, "deformed_beta": deformed}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global LR
    args = parse()
    cfg = CONFIGS[args.config]
    LR = cfg.get("lr", LR)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
