"""Synthetic worm-video generator and dataset (input side of the hot path).

Re-statement of the semantics of WUtils/Simulator.py::generate_video (:20-77) with the separable
structure of its Gaussian cells exploited: video_t = sum_k traces[k,t] * gx_k (x) gy_k (x) gz_k
instead of K*T scipy pdf evaluations over all voxels (O(T*K*N) host work in the reference, unusable
beyond the demo size).  On a GPU the clean frames come from the library's generator kernel
(`dnmf_render_cells`, csrc/dnmf_aux.cu); on the CPU (tests, the reference arm's workload) from one
batched GEMM per chunk of frames.

  * cells:   exp(-|p - P_k(t)|^2 / (2*shape_std))            Simulator.py:72,197-203 (cov = shape_std*I)
  * traces:  1 + Bernoulli(density) spikes (*) exp(-0.3 j), j < 10      Simulator.py:174-195
  * motion:  'gp' -- centres U(0,1)*sz plus, per axis and per frame, an independent draw of a GP prior
             with kernel sigma_d * RBF(ls_d) over the centre coordinate of that axis    Simulator.py:362-391
  * video /= sum(video^2); += 10^(bg_snr/20) * N(0,1); /= max       Simulator.py:66-77
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset


def exponential_traces(K: int, T: int, density: float = .1, b: float = 1.0,
                       rng: Optional[np.random.Generator] = None) -> np.ndarray:
    rng = rng or np.random.default_rng()
    kernel = np.exp(np.arange(0, -3, -.3))
    n = T + len(kernel) - 1
    traces = np.full((K, T), float(b))
    nnz = int(round(density * n))
    for k in range(K):
        spikes = np.zeros(n)
        spikes[rng.choice(n, size=nnz, replace=False)] = 1.0     # scipy.sparse.rand places exactly density*n
        traces[k] += np.convolve(spikes, kernel, "valid")
    return traces


def gp_motion(K: int, T: int, sigma: Sequence[float], ls: Sequence[float], sz: Sequence[int],
              rng: Optional[np.random.Generator] = None) -> torch.Tensor:
    """positions[K,3,T] float32."""
    rng = rng or np.random.default_rng()
    centres = rng.random((K, 3)) * np.asarray(sz, float)
    pos = np.zeros((K, 3, T))
    for d in range(3):
        a = centres[:, d]
        cov = sigma[d] * np.exp(-.5 * (a[:, None] - a[None, :]) ** 2 / ls[d] ** 2)
        w, v = np.linalg.eigh(cov)
        root = v * np.sqrt(np.clip(w, 0, None))[None, :]
        pos[:, d, :] = a[:, None] + root @ rng.standard_normal((K, T))
    return torch.tensor(pos).float()


def quadratic_motion(K: int, T: int, sz: Sequence[int], scale: float = 1.0,
                     rng: Optional[np.random.Generator] = None) -> torch.Tensor:
    """positions[K,3,T] from a random per-frame quadratic map of fixed centres (smooth in space)."""
    rng = rng or np.random.default_rng()
    size = np.asarray(sz, float)
    centres = rng.random((K, 3)) * size
    c = centres / np.maximum(size - 1, 1)                      # normalised so coefficients are in pixels
    phi = np.concatenate((np.ones((K, 1)), c, c * c, c[:, [0]] * c[:, [1]], c[:, [0]] * c[:, [2]],
                          c[:, [1]] * c[:, [2]]), 1)           # [K,10]
    amp = scale * np.array([1.0, 1.0, 0.05])[None, :, None]
    coef = rng.standard_normal((10, 3, T)) * amp
    disp = np.einsum("ka,adt->kdt", phi, coef)
    return torch.tensor(centres[:, :, None] + disp).float()


def render_clean(positions: torch.Tensor, traces, sz: Sequence[int], shape_std: float, device=None,
                 chunk: int = 32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Noise-free frames [T,X,Y,Z] = sum_k traces[k,t] * exp(-|p-P_k(t)|^2/(2*shape_std))."""
    device = torch.device(device) if device is not None else positions.device
    X, Y, Z = (int(s) for s in sz)
    K, _, T = positions.shape
    if device.type == "cuda" and Z <= 64:
        # on the GPU the frames come from the hand-written generator kernel (dnmf_render_cells)
        from .engine import render_cells
        return render_cells(positions, traces, (X, Y, Z), shape_std, device=device, out=out)
    tr = torch.as_tensor(np.asarray(traces), dtype=torch.float32, device=device)
    pos = positions.to(device)
    ax = [torch.arange(n, device=device, dtype=torch.float32) for n in (X, Y, Z)]
    if out is None:
        out = torch.empty(T, X, Y, Z, device=device)
    for t0 in range(0, T, chunk):
        p = pos[:, :, t0:t0 + chunk].permute(2, 0, 1)                      # [t,K,3]
        g = [torch.exp(-(ax[d][None, None, :] - p[:, :, d, None]) ** 2 / (2 * shape_std)) for d in range(3)]
        gx = g[0] * tr[:, t0:t0 + chunk].T[:, :, None]                    # [t,K,X]
        gyz = (g[1][:, :, :, None] * g[2][:, :, None, :]).reshape(p.shape[0], K, Y * Z)
        out[t0:t0 + chunk] = torch.bmm(gx.transpose(1, 2), gyz).reshape(-1, X, Y, Z)
    return out


def generate_video(K, T, sz=(20, 20, 1), shape_std=3, density=.1, bg_snr=-1, traces="exp", motion="gp",
                   motion_par: Optional[Dict] = None, seed: Optional[int] = None, device=None,
                   frame_major: bool = False) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """Same return convention as Simulator.generate_video: (video[X,Y,Z,T], positions[K,3,T], traces[K,T]);
    frame_major=True returns video as [T,X,Y,Z] without the final permute (no extra copy at scale)."""
    rng = np.random.default_rng(seed)
    size = [int(s) for s in (sz.tolist() if torch.is_tensor(sz) else sz)]
    motion_par = motion_par or {}
    if motion == "gp":
        positions = gp_motion(K, T, motion_par.get("sigma", [5, 5, .01]), motion_par.get("ls", [10, 10, 10]), size, rng)
    elif motion == "quadratic":
        positions = quadratic_motion(K, T, size, motion_par.get("scale", 1.0), rng)
    elif motion == "static":
        positions = torch.tensor(rng.random((K, 3)) * np.asarray(size, float)).float()[:, :, None].repeat(1, 1, T)
    else:
        raise ValueError("motion must be 'gp', 'quadratic' or 'static'")
    if isinstance(traces, str):
        if traces != "exp":
            raise ValueError("traces must be 'exp' or an array")
        traces = exponential_traces(K, T, density, rng=rng)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    video = render_clean(positions, traces, size, shape_std, device=device)
    gen = torch.Generator(device=video.device)
    gen.manual_seed(int(rng.integers(0, 2 ** 31 - 1)))
    bg_std = float(np.sqrt(10 ** (bg_snr / 10)))
    chunk = 64                                   # chunked passes: no second copy of the slab (110 GB at T = 40k)
    ssq = sum((video[t0:t0 + chunk].double() ** 2).sum() for t0 in range(0, T, chunk))
    ssq = ssq.float()
    vmax = torch.zeros((), device=video.device)
    for t0 in range(0, T, chunk):
        v = video[t0:t0 + chunk]
        v /= ssq
        v += bg_std * torch.randn(v.shape, generator=gen, device=video.device)
        vmax = torch.maximum(vmax, v.max())
    for t0 in range(0, T, chunk):
        video[t0:t0 + chunk] /= vmax
    if not frame_major:
        video = video.permute(1, 2, 3, 0)
    return video, positions, traces


class SimulatedVideoDataset(Dataset):
    """Drop-in for Demix/dNMF.py:196-217: items are (frame[X,Y,Z] clamped at 0, idx).
    `.video` is [X,Y,Z,T] like the reference; frames are kept frame-major internally."""
    returns_frame_index = True   # item = (frame, its own index): loaders over an attached video are walked for ids only

    def __init__(self, K, T, sz, shape_std, density, bg_snr, traces, motion, motion_par, seed=None, device="cpu"):
        frames, positions, tr = generate_video(K, T, sz, shape_std, density, bg_snr, traces, motion, motion_par,
                                               seed=seed, device=device, frame_major=True)
        self.frames = frames.float().clamp_(min=0).contiguous()       # clamp of Demix/dNMF.py:215, done once
        self.positions = positions
        self.traces = tr

    @property
    def video(self) -> torch.Tensor:
        return self.frames.permute(1, 2, 3, 0)

    def __len__(self):
        return self.frames.shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return self.frames[idx], idx


class FrameDataset(Dataset):
    """Wraps existing frames [T,X,Y,Z] (e.g. a golden fixture or a rank's slab) as (frame, id) items.  `offset` is
    the global id of the slab's first frame: items carry GLOBAL ids (idx + offset), which a model built with
    `frame_offset=offset` maps back to its slab."""
    returns_frame_index = True

    def __init__(self, frames: torch.Tensor, offset: int = 0):
        self.frames = frames
        self.offset = int(offset)

    def __len__(self):
        return self.frames.shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return self.frames[idx], idx + self.offset


class NeuroPALVideoDataset(Dataset):
    """Real-data loader with the semantics of Demix/dNMF.py:220-248 and portable paths: `file` is a directory
    holding data.mat (variable `data` [X,Y,Z,T]) and traces_n.mat (`positions` [K,3,T] 1-based,
    `neuron_names`).  The reference subsamples [::2, ::2, ::10, :100] and rescales the positions to match;
    the strides and frame count are arguments here with the same defaults."""
    returns_frame_index = True

    def __init__(self, file, stride=(2, 2, 10), frames=100):
        import os
        from scipy.io import loadmat
        vid = loadmat(os.path.join(file, "data.mat"))["data"]
        sx, sy, sz_ = stride
        video = np.ascontiguousarray(vid[::sx, ::sy, ::sz_, :frames]).astype(np.float32)
        self.frames = torch.from_numpy(np.ascontiguousarray(np.moveaxis(video, 3, 0))).clamp_(min=0)
        pos_mat = loadmat(os.path.join(file, "traces_n.mat"))
        positions = torch.tensor(np.asarray(pos_mat["positions"], dtype=np.float32)) - 1
        positions[:, 0, :] /= sx
        positions[:, 1, :] /= sy
        positions[:, 2, :] /= sz_
        self.positions = positions
        self.names = pos_mat["neuron_names"][0] if "neuron_names" in pos_mat else None

    @property
    def video(self) -> torch.Tensor:
        return self.frames.permute(1, 2, 3, 0)

    def __len__(self):
        return self.frames.shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return self.frames[idx], idx
