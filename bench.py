#!/usr/bin/env python
"""Benchmark of the dNMF fit hot path (BASELINE.json metric: frame-iterations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2]

One "step" = one Adam iteration of DeformableNMF.update_motion over a batch of B frames
(default B = all T frames of the rank's slab); one frame-iteration = forward +
loss + beta-gradient + Adam for one frame.  N > 1 is launched by torchrun (one rank per GPU); frames
are sharded (weak scaling: every rank owns T frames, every rank the SAME synthetic slab so that the ranks
differ by their hardware only) and the only collective is an 8-byte loss all-reduce every 10 steps.

The JSON line carries, next to the contract's keys:
  value / e2e / roofline / cpu_baseline / clocks     the headline configuration (cfg2 unless --config)
  deformed_beta       the same fused pass late in a fit (every frame its own deformation), with its own frac
  reference_batch     the reference's minibatch size (4 frames) through update_motion (dnmf_motion_epoch)
  e2e_reference_batch the same from HOST frames (4-frame steps: what --impl reference steps)
  e2e_resident        host frames uploaded once (attach_video) + the same steps from ids (public API, amortised)
  trace_update        dnmf_mu_stats + 50 sweeps (update_footprints), frame-MU-iterations/s
  shared_learning     EXTENSION: iterations that also learn positions / widths / background (one all-reduce per step)
  configs             compact legs of the other BASELINE configurations (cfg3, cfg4): value, frac, trace update
  scaling_detail      N > 1: per-rank step times, list lengths, time inside the final loss reduction
  strong_scaling      N > 1: a fixed T_global split over the ranks (BASELINE configuration 5)

`--impl reference` times the reference's CPU torch path (the oracle's TorchPort, which is
bit-identical to /root/reference's code -- tests/test_oracle.py) on the host cores, on the same
workload shape.  It is the only place besides cpu_baseline where bench.py executes oracle/.
"""
import argparse
import ctypes
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
    "cfg3": dict(sz=(512, 256, 32), K=300, T=625, T_single=200, sigma=3.0, shape_std=3.0, deformation="quadratic",
                 lr=1e-7, desc="whole-brain 512x256x32, K=300, T=5000 over 8 GPUs (625 per GPU), quadratic"),
    "cfg4": dict(sz=(256, 128, 21), K=1000, T=250, T_single=100, sigma=6.0, shape_std=18.0, deformation="quadratic",
                 lr=1e-7, desc="dense stress 256x128x21, K=1000, sigma=6, T=2000 over 8 GPUs (250 per GPU)"),
}
CUTOFF = 3.5
LR = 1e-5   # demo.py:42; the raw-pixel quadratic basis needs a smaller step on large volumes (one step of 1e-5
            # moves a voxel at x=511 by 2.6 px through the x^2 term) -> per-config "lr"
MOTION = {"sigma": [5, 5, .01], "ls": [10, 10, 10]}
try:
    ORIG_AFFINITY = os.sched_getaffinity(0)
except Exception:   # pragma: no cover
    ORIG_AFFINITY = None


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
    ap.add_argument("--no-legs", action="store_true", help="skip the cfg3 / cfg4 / strong-scaling legs")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per reference step (default 4)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks, host topology
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


def bind_to_gpu_numa_node(local):
    """Pin this process (and with it the first-touch placement of its pinned host slab) to the cores of the
    NUMA node the GPU hangs off.  Returns a small record for the JSON line."""
    rec = {"numa_node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        rec["numa_node"] = node
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                rec["cpus"] = len(allowed)
    except Exception as ex:
        rec["error"] = repr(ex)[:120]
    return rec


# ------------------------------------------------------------------------------------------------
# reference / CPU arm (oracle port of the reference's torch path)
# ------------------------------------------------------------------------------------------------
def make_cpu_workload(cfg, frames, seed=0):
    from dnmf_b200.simulate import generate_video
    vid, positions, _ = generate_video(cfg["K"], frames, cfg["sz"], cfg["shape_std"], .2, -120, "exp", "gp", MOTION,
                                       seed=seed, device="cpu", frame_major=True)
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
            "config": {"workload": cfg["desc"], "name": args.config, "frames_per_step": B, "lr": LR,
                       "note": "the unit is per frame; the CUDA arm's own 4-frame-step numbers are its "
                               "`reference_batch` (resident) and `e2e_reference_batch` (host frames) entries"},
            "cpu_baseline": {"value": value, "unit": "frame-iterations/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "frame-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm: helpers
# ------------------------------------------------------------------------------------------------
class Loader:
    def __init__(self, items):
        self.items = items

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def build_model(cfg, T, dev, world, tiling=None, seed=100):
    """Synthetic slab generated on the GPU (frame-major), resident in the context.  Every rank generates the SAME
    slab (seed fixed): the ranks then differ by their hardware only, which is what a scaling number should see."""
    from dnmf_b200 import DeformableNMF
    from dnmf_b200.simulate import generate_video
    vid, positions, _ = generate_video(cfg["K"], T, cfg["sz"], cfg["shape_std"], .2, -120, "exp", "gp", MOTION,
                                       seed=seed, device=dev, frame_major=True)
    vid.clamp_(min=0)
    torch.manual_seed(1234)
    dn = DeformableNMF(cfg["sz"], cfg["K"], T, positions=positions[:, :, 0], cutoff=CUTOFF,
                       deformation=cfg["deformation"], shape_std=cfg["sigma"], device=dev, tiling=tiling, verbose=False,
                       global_batch_scale=world)
    dn.attach_video(vid, layout="TXYZ", copy=False)     # the context reads the generated slab in place
    return dn, vid


def list_stats(eng, sz, T, dev):
    """(k_eff, listed): in-cutoff (voxel, neuron) pairs per voxel from the device ranges, and the pairs the tile
    lists hold at the identity deformation."""
    N = int(np.prod(sz))
    rng = eng.ranges()
    ext = np.minimum(rng[:, :, 1] - rng[:, :, 0] + 2, np.asarray(sz)[None, :]).clip(min=0)
    k_eff = float(ext.prod(1).sum()) / N
    beta_id = torch.zeros(10, 3, T, device=dev)
    beta_id[1, 0], beta_id[2, 1], beta_id[3, 2] = 1.0, 1.0, 1.0
    counts, _, _, _ = eng.bin_tiles(beta_id, torch.zeros(1, dtype=torch.int32, device=dev))
    tl = eng.tiling()
    listed = float((counts.astype(np.float64) * tl["tx"] * tl["ty"] * tl["tz"]).sum()) / N
    return k_eff, listed, tl


def time_fit_kernel(eng, dn, beta, ids, reps, tag="fit"):
    """fit_tile_kernel + its second-stage reduction (dnmf_loss_grad), ms per launch, CUDA events.
    DNMF_PROFILE_RANGE=<tag> brackets one launch with cudaProfilerStart/Stop (ncu --profile-from-start off)."""
    from dnmf_b200 import _lib
    B = int(ids.numel())
    dev = beta.device
    grad = torch.zeros(10, 3, eng.T, dtype=torch.float32, device=dev)
    sse = torch.zeros(B, dtype=torch.float64, device=dev)
    lib = eng.lib

    def fit_only():
        _lib.check(lib.dnmf_loss_grad(eng._h, None, ctypes.c_void_p(ids.data_ptr()), B, B,
                                      ctypes.c_void_p(beta.data_ptr()), ctypes.c_void_p(dn.C.data_ptr()),
                                      ctypes.c_void_p(grad.data_ptr()), ctypes.c_void_p(sse.data_ptr()),
                                      eng.stream), "dnmf_loss_grad")
    for _ in range(3):
        fit_only()
        torch.cuda.synchronize()       # lets the library's per-launch choice of the main-loop variant settle
    if os.environ.get("DNMF_PROFILE_RANGE") == tag:
        torch.cuda.cudart().cudaProfilerStart()
        fit_only()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        fit_only()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps, fit_only


def deform(beta, affine, T, dev):
    """Every frame its own translation / linear / quadratic terms (translations sigma = 2 px in x, y and 0.5 px in z,
    linear terms 0.5 %): the late-fit state, no two frames alike.  Returns the identity copy to restore."""
    gen = torch.Generator().manual_seed(7)
    scale = torch.tensor([2.0, 5e-3, 5e-3, 5e-3, 1e-5, 1e-5, 1e-5, 1e-5, 1e-5, 1e-5])[:, None, None]
    if affine:
        scale[4:] = 0.0
    scale = scale * torch.tensor([1.0, 1.0, 0.25])[None, :, None]   # z is shallow: a quarter of the x, y motion
    keep = beta.clone()
    beta.add_((scale * torch.randn(10, 3, T, generator=gen)).to(dev))
    return keep


def time_trace_update(eng, dn, beta, ids_all, T, chunk=250, iters=50):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(0, T, chunk):                  # warm-up at full size: scratch allocations, lazy module load
        eng.mu_stats(ids_all[i:i + chunk], beta)
    eng.mu_sweeps(dn.C.clone(), 0.0, 1)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(0, T, chunk):
        eng.mu_stats(ids_all[i:i + chunk], beta)
    ev1.record()
    torch.cuda.synchronize()
    stats_ms = ev0.elapsed_time(ev1)
    Cmu = dn.C.clone()
    ev0.record()
    eng.mu_sweeps(Cmu, 0.0, iters)
    ev1.record()
    torch.cuda.synchronize()
    sweeps_ms = ev0.elapsed_time(ev1)
    path = eng.mu_path()
    return {"metric": "frame-MU-iterations/s", "value": T * iters / ((stats_ms + sweeps_ms) * 1e-3), "frames": T,
            "iter_c": iters, "stats_ms": stats_ms, "sweeps_ms": sweeps_ms,
            "stats_kernel": "fused tiles" if path & 1 else ("tensor-core panel (tcgen05 tf32)" if path & 4 else "SIMT panel"),
            "what": "dnmf_mu_stats over all frames + %d multiplicative sweeps (update_footprints without the dense "
                    "returns); reference: 1 035 frame-MU-iters/s on the 8-core CPU at cfg1" % iters}


def shared_learning_leg(cfg, dev, world, dist, steps, frames=250):
    """EXTENSION leg (no reference counterpart, flag OFF by default): full-batch iterations that also learn the shared
    parameters -- positions, widths, background -- through DeformableNMF.update_motion after enable_shared_learning:
    fused kernel with the residual written, parameter-gradient kernel, ONE all-reduce of the packed shared gradients
    over the ranks (NCCL), device Adam on beta and on pos / sigma / b, tables and candidate lists rebuilt by kernels."""
    T = min(frames, cfg["T"])
    dn, vid = build_model(cfg, T, dev, world)
    dn.enable_shared_learning(lr_pos=1e-4, lr_sigma=1e-5, lr_background=1e-5)
    opt = torch.optim.Adam([dn.fp.beta], lr=cfg.get("lr", LR))
    ids = torch.arange(T, dtype=torch.int32)
    dn.update_motion(Loader([(None, ids)] * 2), opt, epochs=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    dn.update_motion(Loader([(None, ids)] * steps), opt, epochs=1)
    ev1.record()
    torch.cuda.synchronize()
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    moved = float((dn.fp.pos.detach().cpu() - dn._positions).abs().max())
    out = {"value": T * steps * world / (float(t_ms) * 1e-3), "unit": "frame-iterations/s", "frames_per_gpu": T,
           "steps": steps, "ms_per_step": float(t_ms) / steps, "collective": "all_reduce of 4K+2 doubles per iteration"
           if world > 1 else "none (one rank)", "max_position_change_px": moved,
           "what": "EXTENSION (no reference counterpart): update_motion after enable_shared_learning -- beta, positions, "
                   "widths and background learned together, device-resident step (dnmf_ext_step_begin / _end)"}
    dn.fp.engine.close()
    del dn, vid
    torch.cuda.empty_cache()
    return out


def fp32_peak(eng, local):
    from dnmf_b200 import _lib
    peak = ctypes.c_double(0.0)
    _lib.check(eng.lib.dnmf_measure_fp32_peak(local, 5, ctypes.byref(peak)), "dnmf_measure_fp32_peak")
    return peak.value


def config_leg(name, dev, local, world, dist, steps, warmup, peak):
    """Compact leg of another BASELINE configuration on this rank's GPU: device-resident full-batch steps (value,
    max over ranks), the fused kernel alone (frac), the same with a deformation per frame, the trace update."""
    cfg = CONFIGS[name]
    lr = cfg.get("lr", 1e-5)
    T = cfg["T"] if world > 1 else cfg.get("T_single", cfg["T"])
    sz, K = cfg["sz"], cfg["K"]
    N = int(np.prod(sz))
    dn, vid = build_model(cfg, T, dev, world)
    del vid
    torch.cuda.empty_cache()
    eng = dn.fp.engine
    opt = torch.optim.Adam([dn.fp.beta], lr=lr)
    _, st = dn._adam_state(opt)
    beta = dn.fp.beta.detach()
    ids_all = torch.arange(T, dtype=torch.int32, device=dev)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    step_no = 0

    def one_step():
        nonlocal step_no
        step_no += 1
        eng.motion_step(ids_all, beta, st["exp_avg"], st["exp_avg_sq"], dn.C, lr, (0.9, 0.999), 1e-8, step_no,
                        dn.affine, frames=None, B_global=T * world, loss_out=loss)
    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_step()
    ev1.record()
    torch.cuda.synchronize()
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    value = T * steps * world / (float(t_ms) * 1e-3)
    # restore the identity for the kernel-only numbers (a handful of Adam steps at lr 1e-7 is still the early-fit state)
    k_eff, listed, tl = list_stats(eng, sz, T, dev)
    fit_ms, _ = time_fit_kernel(eng, dn, beta, ids_all, max(3, steps), tag=name)
    flops = N * (144.0 + 25.0 * k_eff)
    frac = flops * T / (fit_ms * 1e-3) * 1e-12 / peak
    keep = deform(beta, dn.affine, T, dev)
    d_ms, _ = time_fit_kernel(eng, dn, beta, ids_all, max(3, steps), tag=name + "_deformed")
    beta.copy_(keep)
    mu = time_trace_update(eng, dn, beta, ids_all, T, chunk=min(T, 250))
    out = {"workload": cfg["desc"], "frames_per_gpu": T, "n_gpus": world, "lr": lr,
           "value": value, "unit": "frame-iterations/s", "ms_per_step": float(t_ms) / steps,
           "roofline": {"bound": "fp32", "frac": frac, "achieved": flops * T / (fit_ms * 1e-3) * 1e-12, "peak": peak,
                        "unit": "TFLOP/s", "kernel_ms_per_launch": fit_ms, "frames_per_launch": T,
                        "flops_per_frame_iter": flops, "k_eff_in_cutoff_pairs_per_voxel": k_eff,
                        "listed_pairs_per_voxel": listed,
                        "hbm_frac": 4.0 * N * T / (fit_ms * 1e-3) * 1e-9 / hbm_peak()[0]},
           "deformed_beta": {"value": T / (d_ms * 1e-3), "kernel_ms_per_launch": d_ms,
                             "frac": flops * T / (d_ms * 1e-3) * 1e-12 / peak},
           "trace_update": mu, "tiling": tl}
    eng.close()
    del dn
    torch.cuda.empty_cache()
    return out


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def h2d_probe(dev, world, dist, nbytes=1 << 30):
    """All ranks copy `nbytes` of pinned host memory to their GPU at the same time: the host-link ceiling the
    end-to-end number runs against (GB/s per GPU, min over ranks)."""
    src = torch.empty(nbytes // 4, dtype=torch.float32, pin_memory=True)
    src.zero_()
    dst = torch.empty_like(src, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    g = torch.tensor([3 * nbytes / (ev0.elapsed_time(ev1) * 1e-3) * 1e-9], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
    del src, dst
    return float(g)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import torch.distributed as dist
    from dnmf_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)      # before the pinned slab is allocated (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sz, K = cfg["sz"], cfg["K"]
    T = args.frames or cfg["T"]
    B = args.batch or T
    N = int(np.prod(sz))
    tiling = tuple(int(v) for v in args.tiling.split(",")) if args.tiling else None
    dn, vid = build_model(cfg, T, dev, world, tiling)
    eng = dn.fp.engine
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
    ev0, ev1, evk = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ev0.record()
    for i in range(args.steps):
        one_step(args.warmup + i)
    evk.record()                           # the rank's own K steps are enqueued up to here
    drain(final=True)                      # the timed region ends after the last loss reduction
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ms_steps_only = ev0.elapsed_time(evk)
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
    frame_iters = B * args.steps * world
    value = frame_iters / (float(t_ms) * 1e-3)
    ms_own = ms_total

    # ---- per-rank detail of the scaling number (N > 1): who is slow, and is it the data, the GPU or the reduction ----
    k_eff, listed, tl = list_stats(eng, sz, T, dev)
    scaling_detail = None
    if world > 1:
        try:
            clk = float(torch.cuda.clock_rate(dev))
        except Exception:
            clk = 0.0
        mine = torch.tensor([ms_own / args.steps, ms_steps_only / args.steps, ms_own - ms_steps_only, listed, clk],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        if rank == 0:
            rows = torch.stack(allr).cpu().numpy()
            scaling_detail = {"per_rank_ms_per_step": rows[:, 0].tolist(),
                              "per_rank_ms_per_step_kernels_only": rows[:, 1].tolist(),
                              "per_rank_device_ms_from_last_step_to_end_of_final_loss_reduction": rows[:, 2].tolist(),
                              "per_rank_listed_pairs_per_voxel": rows[:, 3].tolist(),
                              "per_rank_sm_clock_mhz_after": rows[:, 4].tolist(),
                              "workloads": "identical on every rank (same seed): differences are the hardware's",
                              "note": "ms_per_step = CUDA events around the K steps + the final loss all-reduce on "
                                      "this rank; kernels_only stops before the reduction; `value` uses the MAX over ranks"}

    # ---- strong scaling (BASELINE configuration 5): a fixed T_global split over the ranks ----
    strong = None
    if world > 1 and not args.no_legs:
        from dnmf_b200.sharding import frame_slab
        strong = []
        for Tg in (T,):
            _, cnt = frame_slab(Tg, world, rank)
            ids_s = ids_all[:cnt]
            sno = step_no[0]

            def s_step():
                nonlocal sno
                sno += 1
                eng.motion_step(ids_s, beta, st["exp_avg"], st["exp_avg_sq"], dn.C, LR, (0.9, 0.999), 1e-8, sno,
                                dn.affine, frames=None, B_global=Tg, loss_out=loss_ring[0:1])
            for _ in range(3):
                s_step()
            barrier()
            ev0.record()
            for _ in range(args.steps):
                s_step()
            ev1.record()
            torch.cuda.synchronize()
            ts = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            strong.append({"T_global": Tg, "frames_per_gpu": cnt, "n_gpus": world,
                           "value": Tg * args.steps / (float(ts) * 1e-3), "ms_per_step": float(ts) / args.steps,
                           "unit": "frame-iterations/s",
                           "note": "total work fixed as N grows (the N = 1 run's headline `value` is the same "
                                   "T_global on one GPU); tools/sweep.sh produces the T = 1k..40k table of BASELINE "
                                   "configuration 5"})
            step_no[0] = sno

    # ---- end to end: public API, host (pinned) frames, H2D + loss read-back inside the timed region ----
    e2e = e2e_b4 = e2e_res = None
    if host_frames is not None:
        link = h2d_probe(dev, world, dist)
        Be = min(B, T)
        host_batches = [(host_frames[i:i + Be], torch.arange(i, i + Be, dtype=torch.int32))
                        for i in range(0, T - Be + 1, Be)]
        dn._video_resident = False
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
        gbps = (Be * N * 4 + Be * 4) * steps_e / (float(t_e) * 1e-3) * 1e-9
        e2e = {"value": Be * steps_e * world / (float(t_e) * 1e-3), "unit": "frame-iterations/s",
               "h2d_bytes_per_step": Be * N * 4 + Be * 4, "d2h_bytes_per_step": 8,
               "h2d_gbps_per_gpu": gbps, "h2d_link_ceiling_gbps_per_gpu": link,
               "fraction_of_link_ceiling": gbps / link if link else None,
               "bound": "host-to-device link: every step ships its %.2f GB of fp32 frames from pinned host memory in "
                        "96 MB chunks on a copy stream, double-buffered against the fused kernel.  The ceiling is "
                        "measured in this run: all ranks copying 1 GB of pinned memory at the same time "
                        "(h2d_link_ceiling_gbps_per_gpu, min over ranks)" % (Be * N * 4 / 1e9),
               "host": numa,
               "api": "DeformableNMF.update_motion(host loader, torch.optim.Adam) -> dnmf_motion_step_host"}
        # the reference's own step: 4 host frames per Adam step (what `--impl reference` steps)
        if T >= 8 and rank == 0:
            Bq = 4
            items = [(host_frames[i:i + Bq], torch.arange(i, i + Bq, dtype=torch.int32)) for i in range(0, min(T, 400) - Bq + 1, Bq)]
            dn.update_motion(Loader(items[:4]), opt, epochs=1)
            torch.cuda.synchronize()
            ev0.record()
            dn.update_motion(Loader(items), opt, epochs=1)
            ev1.record()
            torch.cuda.synchronize()
            e2e_b4 = {"batch": Bq, "steps": len(items), "unit": "frame-iterations/s",
                      "value": len(items) * Bq / (ev0.elapsed_time(ev1) * 1e-3),
                      "h2d_bytes_per_step": Bq * N * 4 + Bq * 4, "d2h_bytes_per_step": 8,
                      "api": "DeformableNMF.update_motion(4-frame host minibatches) -> dnmf_motion_step_host; "
                             "the same batch size as the reference arm's steps"}
        barrier()
        # frames shipped ONCE (attach_video), then the same steps from ids: how the public API is meant to be used
        # over many epochs.  Timed: the upload + `steps` full-batch steps + one read-back of the losses.
        dn.loss_history.clear()
        id_batches = [(None, torch.arange(i, i + Be, dtype=torch.int32)) for i in range(0, T - Be + 1, Be)]
        id_loader = Loader([id_batches[i % len(id_batches)] for i in range(steps_e)])
        barrier()
        ev0.record()
        dn.attach_video(host_frames, layout="TXYZ")
        dn.update_motion(id_loader, opt, epochs=1)
        _ = dn.losses()
        ev1.record()
        barrier()
        t_r = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
        e2e_res = {"value": Be * steps_e * world / (float(t_r) * 1e-3), "unit": "frame-iterations/s",
                   "h2d_bytes_total": T * N * 4, "steps": steps_e,
                   "what": "attach_video(host frames) once + %d full-batch update_motion steps from frame ids + the "
                           "losses read back; the upload is inside the timed region and amortised over the steps "
                           "(a fit runs 10-50 epochs per update_motion call, demo.py:45)" % steps_e}
        dn._video_resident = True
        del host_frames
        host_frames = None

    # ---- extension: shared-parameter learning (all ranks: it carries the path's one real collective) ----
    shared = None
    if not args.no_legs:
        try:
            shared = shared_learning_leg(cfg, dev, world, dist, max(3, min(args.steps, 10)))
        except Exception as ex:  # pragma: no cover
            shared = {"error": repr(ex)[:300]}
        barrier()

    # ---- roofline of the dominant kernel (fit_tile_kernel), timed live with CUDA events (rank 0) ----
    roofline = deformed = ref_batch = mu = None
    peak = 0.0
    if rank == 0:
        ids = batches[0]
        reps = max(5, args.steps)
        fit_ms, _ = time_fit_kernel(eng, dn, beta, ids, reps)
        flops_per_frame = N * (144.0 + 25.0 * k_eff)
        bytes_per_frame = 4.0 * N
        peak = fp32_peak(eng, local)
        hbm, hbm_src = hbm_peak()
        ach_tf = flops_per_frame * B / (fit_ms * 1e-3) * 1e-12
        ach_gbs = bytes_per_frame * B / (fit_ms * 1e-3) * 1e-9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per frame of fit_tile_kernel from the committed `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "fit_tile_traffic.json")))
            if tj.get("config") == args.config:
                traffic, traffic_src = tj["dram_bytes_per_frame"] * B, tj["source"]
        except Exception:
            pass
        roofline = {"bound": "fp32", "achieved": ach_tf, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach_tf / peak if peak else None, "traffic": traffic,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                    "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_frame * B,
                    "kernel": "fit_tile_kernel", "kernel_ms_per_launch": fit_ms, "frames_per_launch": B,
                    "flops_per_frame_iter": flops_per_frame, "k_eff_in_cutoff_pairs_per_voxel": k_eff,
                    "listed_pairs_per_voxel": listed,
                    "peak_source": "dnmf_measure_fp32_peak (FFMA microbenchmark, measured live in this run)",
                    "hbm": {"achieved": ach_gbs, "peak": hbm, "unit": "GB/s", "frac": ach_gbs / hbm,
                            "bytes_per_frame_iter": bytes_per_frame, "peak_source": hbm_src}}

        # ---- the same fused pass late in a fit, when every frame has its own deformation ----
        try:
            keep = deform(beta, dn.affine, T, dev)
            d_ms, _ = time_fit_kernel(eng, dn, beta, ids, reps, tag="deformed")
            beta.copy_(keep)
            deformed = {"value": B / (d_ms * 1e-3), "unit": "frame-iterations/s", "kernel_ms_per_launch": d_ms,
                        "frac": flops_per_frame * B / (d_ms * 1e-3) * 1e-12 / peak if peak else None,
                        "what": "fit_tile_kernel + reduction with a different deformation per frame (translations "
                                "sigma = 2 px in x, y and 0.5 px in z, linear terms 0.5 %, no two frames alike): list "
                                "and slices rebuilt for every frame"}
        except Exception as ex:  # pragma: no cover
            deformed = {"error": repr(ex)}

        # ---- the reference's own minibatch size (demo.py: 4 frames) through the public API, resident video ----
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

        # ---- trace update (update_footprints hot loop #2), frame-MU-iterations/s ----
        if not args.no_mu:
            try:
                mu = time_trace_update(eng, dn, beta, ids_all, T)
            except Exception as exc:                          # keep the headline line even if this leg fails
                mu = {"error": str(exc)}

    # ---- the other BASELINE configurations (cfg3, cfg4) on this rank count ----
    legs = None
    if not args.no_legs:
        eng.close()                      # frees the headline configuration's context (video slab, statistics)
        del dn
        torch.cuda.empty_cache()
        pk = torch.tensor([peak], dtype=torch.float64, device=dev)
        if world > 1:
            dist.broadcast(pk, 0)
        legs = {}
        for name in ("cfg3", "cfg4"):
            if name == args.config:
                continue
            try:
                legs[name] = config_leg(name, dev, local, world, dist, max(3, min(args.steps, 10)), 3, float(pk))
            except Exception as ex:  # pragma: no cover
                legs[name] = {"error": repr(ex)[:300]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, ORIG_AFFINITY)   # the reference gets every host core, not just the GPU's node
        except Exception:
            pass
        Bc = args.cpu_frames or 4
        v, ms, cores = time_reference(cfg, Bc, 2, 1)
        cpu_baseline = {"value": v, "unit": "frame-iterations/s", "cores": cores, "kind": "port",
                        "sample": "2 Adam steps of %d frames of the %s volume (after 1 warm-up) through the "
                                  "oracle's torch port of the reference CPU path" % (Bc, args.config),
                        "ms_per_step": ms}

    # per step: check_ids_kernel + tile_windows_kernel (pre-pass), fit_tile_kernel, reduce_partials_kernel, adam_kernel
    launches = sum(c1[k] - c0[k] for k in ("fit_launches", "reduce_launches", "adam_launches", "prepass_launches"))
    line = {"metric": "frame-iterations/s", "value": value, "unit": "frame-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t_ms) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "name": args.config, "frames_per_gpu": T, "frames_per_step_per_gpu": B,
                       "cutoff_sigma": CUTOFF, "lr": LR, "tiling": tl,
                       "l2": "inputs (%.2f GB per step) larger than L2" % (B * N * 4 / 1e9)},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "final_loss": final_loss, "trace_update": mu,
            "reference_batch": ref_batch, "e2e_reference_batch": e2e_b4, "e2e_resident": e2e_res,
            "deformed_beta": deformed, "shared_learning": shared, "configs": legs, "scaling_detail": scaling_detail, "strong_scaling": strong}
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
