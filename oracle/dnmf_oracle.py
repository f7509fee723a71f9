"""CPU oracle for the dNMF fit hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (dnmf_b200/) never does and fails loudly without its CUDA
library.  Parity status: PINNED by executing the real reference in the build container
(tests/test_oracle.py, oracle/make_golden.py -> tests/golden/*.npz); the
reference itself ships no tests or golden vectors.

Two restatements live here (all file:line citations are relative to /root/reference):

1. `TorchPort` -- the reference's own decomposition, re-stated with the same ATen calls
   (einsum -> grid_sample(trilinear, zeros, align_corners=True) -> einsum -> mse_loss ->
   autograd -> torch.optim.Adam over the dense [10,3,T] tensor; Demix/dNMF.py:19-62,181-194)
   and the fp64 numpy multiplicative update (Demix/dNMF.py:139-149,163-179).  This is the
   "port" that bench.py times as the CPU baseline and that the GPU parity tests compare with.

2. `closed_form_*` -- the separable closed form the CUDA kernels implement (SURVEY.md F1/F2,
   Appendix A), in numpy float32 with the exact coordinate op order, plus the integer binning
   spec (`axis_ranges`, `tile_window`, `bin_tiles`) that the binning kernel must match
   bit for bit, and `adam_dense` (torch _single_tensor_adam formula, F4).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32

# ------------------------------------------------------------------------------------------------
# 1. Torch port of the reference decomposition
# ------------------------------------------------------------------------------------------------


def identity_beta(T: int) -> torch.Tensor:
    """beta[10,3,T]: row 0 zero, rows 1..3 identity, rows 4..9 zero (Demix/dNMF.py:24-26)."""
    b = torch.zeros(10, 3, T)
    b[1, 0], b[2, 1], b[3, 2] = 1.0, 1.0, 1.0
    return b


def voxel_basis(sz: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(grid[X,Y,Z,3], phi[X,Y,Z,10]) with phi = [1,x,y,z,x2,y2,z2,xy,xz,yz]
    (Demix/dNMF.py:22-23,46-51)."""
    X, Y, Z = (int(s) for s in sz)
    gx, gy, gz = torch.meshgrid(torch.arange(X), torch.arange(Y), torch.arange(Z), indexing="ij")
    g = torch.stack((gx, gy, gz), 3).float()
    x, y, z = g[..., 0:1], g[..., 1:2], g[..., 2:3]
    phi = torch.cat((torch.ones_like(x), g, g * g, x * y, x * z, y * z), 3)
    return g, phi


def gaussian_volume(grid: torch.Tensor, pos: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """A[X,Y,Z,K] = exp(sum_d -(p_d - pos_kd)^2 / sigma_k^2)  (Demix/dNMF.py:39-40)."""
    d = grid[:, :, :, :, None] - pos.T[None, None, None, :, :]
    return torch.exp((-(d ** 2) / sigma[None, None, None, None, :] ** 2).sum(3))


class TorchPort:
    """Re-statement of ExponentialFP + DeformableNMF (Demix/dNMF.py:18-194) on CPU torch.

    Layouts are the reference's: beta[10,3,T], C[K,T], frames[B,X,Y,Z]."""

    def __init__(self, sz, K, T, positions=None, shape_std=3.0, C0=None):
        self.sz = torch.as_tensor(sz).long()
        self.K, self.T = int(K), int(T)
        self.grid_id, self.phi = voxel_basis(self.sz.tolist())
        self.beta = identity_beta(T).requires_grad_(True)
        self.sigma = torch.ones(K) * shape_std
        if positions is None:
            self.pos = 1 + torch.rand(K, 3) * self.sz[None, :]
        else:
            self.pos = torch.as_tensor(positions).float()
        self.A = gaussian_volume(self.grid_id, self.pos, self.sigma)
        self.C = torch.rand(K, T) if C0 is None else torch.as_tensor(C0).float().clone()

    # -- Demix/dNMF.py:53-58 (reg omitted: detached constant) -------------------------
    def forward(self, times: Sequence[int], C: Optional[torch.Tensor] = None):
        C = self.C if C is None else C
        times = list(times)
        q = torch.einsum("mnza,abt->mnzbt", self.phi, self.beta[:, :, times])
        u = 2 * q / (self.sz[None, None, None, :, None] - 1) - 1
        vol = self.A.permute(3, 2, 1, 0)[None].expand(len(times), -1, -1, -1, -1)
        A_t = F.grid_sample(vol, u.permute(4, 2, 1, 0, 3), mode="bilinear", padding_mode="zeros",
                            align_corners=True).permute(0, 1, 4, 3, 2)
        A_tC = torch.einsum("tkmnz,kt->tmnz", A_t, C[:, times])
        return A_tC, A_t, u

    # -- Demix/dNMF.py:185-191 --------------------------------------------------------------------
    def motion_step(self, frames: torch.Tensor, times: Sequence[int], optimizer,
                    affine: bool = False, reg_gamma: Optional[float] = None) -> float:
        """reg_gamma=None is the reference (detached regulariser, F3).  A number adds the OPT-IN
        differentiable, index-consistent penalty gamma * mean_t [logdetJ_t(sz-1)^2 + logdetJ_t(0)^2]."""
        optimizer.zero_grad()
        A_tC, _, _ = self.forward(times)
        recon = F.mse_loss(A_tC, frames)
        loss = recon
        if reg_gamma is not None:
            b = self.beta[:, :, list(times)]
            hi = (self.sz - 1).float()
            reg = log_det_jac_consistent(b, hi) ** 2 + log_det_jac_consistent(b, torch.zeros(3)) ** 2
            loss = recon + reg_gamma * reg.mean()
        loss.backward()
        if affine:  # "affine" = quadratic rows frozen
            self.beta.grad[4:] = 0
        optimizer.step()
        return float(recon.detach())

    def update_motion(self, batches: Iterable[Tuple[torch.Tensor, Sequence[int]]], optimizer,
                      epochs: int = 1, affine: bool = False) -> List[float]:
        losses = []
        for _ in range(epochs):
            for frames, times in batches:
                losses.append(self.motion_step(frames, list(times), optimizer, affine))
        return losses

    # -- Demix/dNMF.py:69-93 without the NN warp; dense fp64 like the reference -------------------
    def pushforward(self, batches) -> Tuple[np.ndarray, np.ndarray]:
        A_list, Y_list = [], []
        with torch.no_grad():
            for frames, times in batches:
                _, A_t, _ = self.forward(list(times))
                A_list.append(A_t.permute(2, 3, 4, 1, 0).numpy().astype(np.float64))
                Y_list.append(frames.permute(1, 2, 3, 0).numpy().astype(np.float64))
        return np.concatenate(A_list, 4), np.concatenate(Y_list, 3)

    # -- Demix/dNMF.py:163-177 --------------------------------------------------------------------
    def update_footprints(self, batches, gamma_c=1e-2, iter_c=10):
        A_t, Y = self.pushforward(batches)
        C = self.C.numpy()
        Gm, bv = mu_stats_dense(A_t, Y)
        for _ in range(iter_c):
            C = mu_sweep(Gm, bv, C, gamma_c)
        self.C = torch.tensor(C).float()
        return A_t, Y


def mu_stats_dense(A_t: np.ndarray, Y: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """G[k,l,t], b[k,t] in fp64 (Demix/dNMF.py:141-142); hoisting them out of the iteration
    loop is bit-identical to recomputing them each sweep."""
    Gm = np.einsum("mnzkt,mnzlt->klt", A_t, A_t)
    bv = np.einsum("mnzkt,mnzt->kt", A_t, Y)
    return Gm, bv


def mu_sweep(Gm: np.ndarray, bv: np.ndarray, C: np.ndarray, gamma) -> np.ndarray:
    """One multiplicative update of the traces (Demix/dNMF.py:143-148)."""
    C1 = bv.copy()
    C2 = np.einsum("klt,lt->kt", Gm, C)
    if gamma is not None:
        nbr = np.hstack((C[:, 0][:, None], C[:, :-1])) + np.hstack((C[:, 1:], C[:, -1][:, None]))
        C1 = C1 + gamma * nbr
        C2 = C2 + 2 * gamma * C
    return C * C1 / (C2 + 1e-32)


def update_spatial(A: np.ndarray, C: np.ndarray, Y_i: np.ndarray, D=None, gamma=None) -> np.ndarray:
    """Non-parametric footprint update (Demix/dNMF.py:151-160), numpy fp64; the voxel axes of A[..., K] and
    Y_i[..., T] may be any leading shape (the reference's einsum strings fix two)."""
    A = np.asarray(A, np.float64)
    C = np.asarray(C, np.float64)
    Y_i = np.asarray(Y_i, np.float64)
    C_s = np.einsum("kt,pt->kp", C, C)
    A1 = np.einsum("...t,kt->...k", Y_i, C)
    A2 = np.einsum("...k,kp->...p", A, C_s)
    if D is not None:
        A2 = A2 + gamma * np.asarray(D, np.float64)
    return A * A1 / (A2 + 1e-32)


def distance_penalty(sz: Sequence[int], positions: np.ndarray) -> np.ndarray:
    """D = 1 - exp(-0.01 * cdist(grid, positions)) reshaped to [X,Y,Z,K] (Demix/dNMF.py:133-135)."""
    X, Y, Z = (int(s) for s in sz)
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    d = np.sqrt(((g[:, None, :] - np.asarray(positions, np.float64)[None, :, :]) ** 2).sum(-1))
    return (1 - np.exp(-.01 * d)).reshape(X, Y, Z, -1)


def update_temporal_dense(A_t: np.ndarray, C: np.ndarray, Y: np.ndarray, gamma=None) -> np.ndarray:
    """The static update_temporal on dense arrays (Demix/dNMF.py:139-149), numpy fp64."""
    A_t = np.asarray(A_t, np.float64)
    Gm, bv = mu_stats_dense(A_t, np.asarray(Y, np.float64))
    return mu_sweep(Gm, bv, np.asarray(C, np.float64), gamma)


def render_cells(positions: np.ndarray, traces: np.ndarray, sz: Sequence[int], shape_std: float) -> np.ndarray:
    """Clean frames [T,X,Y,Z] of the synthetic generator in fp64: sum_k traces[k,t] * pdf_k / max(pdf_k) with
    pdf_k the normal density of cov = shape_std * I centred at positions[k,:,t] (WUtils/Simulator.py:66-73,197-203:
    the density is rescaled to peak 1, i.e. exp(-|p - P|^2 / (2 shape_std)))."""
    X, Y, Z = (int(s) for s in sz)
    K, _, T = positions.shape
    ax = [np.arange(n, dtype=np.float64) for n in (X, Y, Z)]
    out = np.zeros((T, X, Y, Z))
    for t in range(T):
        for k in range(K):
            g = [np.exp(-(ax[d] - float(positions[k, d, t])) ** 2 / (2.0 * shape_std)) for d in range(3)]
            out[t] += float(traces[k, t]) * g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]
    return out


def log_det_jac_consistent(Bm: torch.Tensor, P) -> torch.Tensor:
    """log|det J| with cross-term rows matching the basis order (xy=7, xz=8, yz=9); opt-in fix of F3."""
    x, y, z = P[0], P[1], P[2]
    rows = []
    for c in range(3):
        rows.append((Bm[1, c] + 2 * Bm[4, c] * x + Bm[7, c] * y + Bm[8, c] * z,
                     Bm[2, c] + 2 * Bm[5, c] * y + Bm[7, c] * x + Bm[9, c] * z,
                     Bm[3, c] + 2 * Bm[6, c] * z + Bm[8, c] * x + Bm[9, c] * y))
    (a, b, c), (d, e, f), (g, h, i) = rows
    return torch.log(abs(a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)))


def log_det_jac(Bm: torch.Tensor, P) -> torch.Tensor:
    """Diagnostic of Demix/dNMF.py:107-122, including its cross-term index order."""
    x, y, z = P[0], P[1], P[2]
    rows = []
    for c in range(3):
        rows.append((Bm[1, c] + 2 * Bm[4, c] * x + Bm[7, c] * y + Bm[9, c] * z,
                     Bm[2, c] + 2 * Bm[5, c] * y + Bm[7, c] * x + Bm[8, c] * z,
                     Bm[3, c] + 2 * Bm[6, c] * z + Bm[8, c] * y + Bm[9, c] * x))
    (a, b, c), (d, e, f), (g, h, i) = rows
    return torch.log(abs(a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)))


# ------------------------------------------------------------------------------------------------
# 2. Closed form (what the CUDA kernels compute), numpy float32
# ------------------------------------------------------------------------------------------------


def axis_ranges(pos: np.ndarray, sigma: np.ndarray, sz: Sequence[int], cutoff: float) -> np.ndarray:
    """Integer node ranges [K,3,2] = (lo, hi) of the truncated footprint per axis.

    lo = max(0, ceil(pos - fl(cutoff*sigma))), hi = min(s-1, floor(pos + fl(cutoff*sigma)));
    cutoff <= 0 means no truncation (lo=0, hi=s-1).  Pure fp32, no FMA: the binning kernel
    must reproduce this bit for bit."""
    pos = np.asarray(pos, f32)
    sigma = np.asarray(sigma, f32)
    K = pos.shape[0]
    out = np.zeros((K, 3, 2), np.int32)
    for d in range(3):
        s = int(sz[d])
        if cutoff <= 0 or not math.isfinite(cutoff):
            out[:, d, 0], out[:, d, 1] = 0, s - 1
            continue
        rad = (f32(cutoff) * sigma).astype(f32)
        lo_f = np.ceil((pos[:, d] - rad).astype(f32))
        hi_f = np.floor((pos[:, d] + rad).astype(f32))
        lo_f = np.fmin(np.fmax(lo_f, f32(0)), f32(s))
        hi_f = np.fmin(np.fmax(hi_f, f32(-1)), f32(s - 1))
        out[:, d, 0] = lo_f.astype(np.int32)
        out[:, d, 1] = hi_f.astype(np.int32)
    return out


def neighbour_lists(rng: np.ndarray) -> List[List[int]]:
    """Static neighbour lists behind the sparse trace sweeps (twin of `mu_build_neighbours` in the CUDA library):
    A_k(ix) != 0 needs lo_k - 1 < ix_d < hi_k + 1 on every axis (table entry i holds (G[i], G[i+1]-G[i]), zero
    outside [lo, hi]), so G_t[k][l] = sum_p A_k A_l (Demix/dNMF.py:141) can only be non-zero when those open
    intervals of k and l meet on all three axes -- whatever the deformation.  The test is the conservative
    lo_k - 1 <= hi_l + 1 and lo_l - 1 <= hi_k + 1; a neuron with an empty range has only itself."""
    K = rng.shape[0]
    out = []
    for k in range(K):
        row = []
        for l in range(K):
            meet = True
            for d in range(3):
                a_lo, a_hi = int(rng[k, d, 0]), int(rng[k, d, 1])
                b_lo, b_hi = int(rng[l, d, 0]), int(rng[l, d, 1])
                meet = meet and a_lo <= a_hi and b_lo <= b_hi and a_lo - 1 <= b_hi + 1 and b_lo - 1 <= a_hi + 1
            if meet or l == k:
                row.append(l)
        out.append(row)
    return out


def axis_tables(pos, sigma, sz, cutoff) -> Tuple[List[np.ndarray], np.ndarray]:
    """Per axis d: table[K, s_d+3, 2] of (G[i], G[i+1]-G[i]) for i = -2..s_d, with
    G[i] = exp(-(i-pos)^2/sigma^2) inside [lo,hi] and 0 outside (zero padding of grid_sample
    + truncation).  Returns (tables, ranges)."""
    pos = np.asarray(pos, f32)
    sigma = np.asarray(sigma, f32)
    rng = axis_ranges(pos, sigma, sz, cutoff)
    tabs = []
    for d in range(3):
        s = int(sz[d])
        i = np.arange(-2, s + 2, dtype=np.int32)               # nodes -2 .. s+1
        delta = (i.astype(f32)[None, :] - pos[:, d][:, None]).astype(f32)
        g = np.exp(-((delta * delta).astype(f32) / (sigma * sigma).astype(f32)[:, None]).astype(f32)).astype(f32)
        live = (i[None, :] >= rng[:, d, 0][:, None]) & (i[None, :] <= rng[:, d, 1][:, None])
        g = np.where(live, g, f32(0)).astype(f32)
        tab = np.stack((g[:, :-1], (g[:, 1:] - g[:, :-1]).astype(f32)), 2)   # i = -2 .. s
        tabs.append(np.ascontiguousarray(tab))
    return tabs, rng


def _fma(a, b, c):
    """fp32 fused multiply-add of float32 arrays: the product of two float32 is exact in float64, so one float64
    addition and one rounding to float32 reproduce the device's fmaf up to double rounding (probability ~2^-29
    per operation)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def deformed_coords(beta_t: np.ndarray, sz: Sequence[int], q_order: str = "sequential") -> List[np.ndarray]:
    """q_d(p) = sum_a phi_a(p) beta_t[a, d] in fp32 for every voxel (Demix/dNMF.py:54).  The reference evaluates
    this contraction with an MKL GEMM whose summation order is not specified; any fp32 order is a legitimate
    realisation and they differ by a few ulp of |q| (3e-5 px at x = 255, 6e-5 px at x = 511).  Orders:
      "sequential"  a = 0..9 with separate multiply and add (the oracle's default),
      "horner"      the CUDA kernels' order: FMA chains over the (x, y) monomials, then Horner in z,
      "exact"       float64 accumulation rounded once to fp32 (the correctly rounded value)."""
    X, Y, Z = (int(s) for s in sz)
    x = np.arange(X, dtype=f32)[:, None, None]
    y = np.arange(Y, dtype=f32)[None, :, None]
    z = np.arange(Z, dtype=f32)[None, None, :]
    one = np.ones((X, Y, Z), f32)
    b = np.asarray(beta_t, f32)
    out = []
    for d in range(3):
        if q_order == "horner":
            v = _fma(b[1, d], x, b[0, d])
            v = _fma(b[2, d], y, v)
            v = _fma(b[4, d], (x * x).astype(f32), v)
            v = _fma(b[5, d], (y * y).astype(f32), v)
            v = _fma(b[7, d], (x * y).astype(f32), v)
            w = _fma(b[9, d], y, _fma(b[8, d], x, b[3, d]))
            q = _fma(z * one, _fma(z * one, b[6, d], w * one), v * one)
        elif q_order == "exact":
            phi = [one, x * one, y * one, z * one, (x * x) * one, (y * y) * one, (z * z) * one, (x * y) * one,
                   (x * z) * one, (y * z) * one]
            q = sum(phi[a].astype(np.float64) * np.float64(b[a, d]) for a in range(10)).astype(f32)
        else:
            phi = [one, x * one, y * one, z * one, (x * x) * one, (y * y) * one, (z * z) * one, (x * y) * one,
                   (x * z) * one, (y * z) * one]
            q = np.zeros((X, Y, Z), f32)
            for a in range(10):
                q = (q + (phi[a] * b[a, d]).astype(f32)).astype(f32)
        out.append(q)
    return out


def sample_coords(beta_t: np.ndarray, sz: Sequence[int], q_order: str = "sequential") -> np.ndarray:
    """Un-normalised sample coordinates ix[3, X, Y, Z] in fp32 with the reference's op order
    (Demix/dNMF.py:54-55 then ATen grid_sampler unnormalize)."""
    X, Y, Z = (int(s) for s in sz)
    qs = deformed_coords(beta_t, sz, q_order)
    out = np.empty((3, X, Y, Z), f32)
    for d in range(3):
        q = qs[d]
        sm1 = f32(int(sz[d]) - 1)
        if int(sz[d]) == 1:
            # The reference divides by s-1 = 0 here (NaN everywhere).  The CUDA path
            # documents a deviation for singleton axes: the sample coordinate is q itself.
            out[d] = q
            continue
        u = (((f32(2) * q).astype(f32) / sm1).astype(f32) - f32(1)).astype(f32)
        out[d] = ((((u + f32(1)).astype(f32)) / f32(2)).astype(f32) * sm1).astype(f32)
    return out


def closed_form_frame(frame: np.ndarray, beta_t: np.ndarray, c_t: np.ndarray, tabs, sz):
    """One frame: returns (yhat[X,Y,Z] f32, sse float64, dsse_dbeta[10,3] float64) where
    sse = sum_p (yhat - Y)^2 and dsse_dbeta = 2 * sum_p phi_a r dYhat/dix_b (no 1/(B N))."""
    X, Y, Z = (int(s) for s in sz)
    ix = sample_coords(beta_t, sz)
    i_idx, fr = [], []
    for d in range(3):
        s = int(sz[d])
        c = np.fmin(np.fmax(ix[d], f32(-2)), f32(s))
        fl = np.floor(c)
        i_idx.append(fl.astype(np.int32) + 2)
        fr.append((c - fl).astype(f32))
    yhat = np.zeros((X, Y, Z), f32)
    g = [np.zeros((X, Y, Z), f32) for _ in range(3)]
    K = tabs[0].shape[0]
    for k in range(K):
        ck = f32(c_t[k])
        a, dd = [], []
        for d in range(3):
            e = tabs[d][k][i_idx[d]]
            a.append((e[..., 0] + fr[d] * e[..., 1]).astype(f32))
            dd.append(e[..., 1])
        yhat += ck * a[0] * a[1] * a[2]
        g[0] += ck * dd[0] * a[1] * a[2]
        g[1] += ck * a[0] * dd[1] * a[2]
        g[2] += ck * a[0] * a[1] * dd[2]
    r = (yhat - frame.astype(f32)).astype(f32)
    sse = float(np.sum(r.astype(np.float64) ** 2))
    x = np.arange(X, dtype=np.float64)[:, None, None]
    y = np.arange(Y, dtype=np.float64)[None, :, None]
    z = np.arange(Z, dtype=np.float64)[None, None, :]
    one = np.ones((X, Y, Z))
    phi = [one, x * one, y * one, z * one, x * x * one, y * y * one, z * z * one, x * y * one,
           x * z * one, y * z * one]
    grad = np.zeros((10, 3))
    for b in range(3):
        h = r.astype(np.float64) * g[b].astype(np.float64)
        for a_ in range(10):
            grad[a_, b] = 2.0 * np.sum(phi[a_] * h)
    return yhat, sse, grad


def closed_form_step(frames: np.ndarray, times: Sequence[int], beta: np.ndarray, C: np.ndarray,
                     tabs, sz) -> Tuple[float, np.ndarray]:
    """MSE loss and dense gradient [10,3,T] (zero off-batch) of one minibatch
    (Demix/dNMF.py:187-190; mean over B*N)."""
    B = len(times)
    N = int(np.prod([int(s) for s in sz]))
    grad = np.zeros(beta.shape, np.float64)
    sse = 0.0
    for j, t in enumerate(times):
        _, s, g = closed_form_frame(frames[j], beta[:, :, t], C[:, t], tabs, sz)
        sse += s
        grad[:, :, t] += g / (B * N)
    return sse / (B * N), grad.astype(f32)


def adam_dense(p, g, m, v, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor formula in fp32 over the WHOLE tensor.
    Returns (p, m, v) as new float32 arrays; `step` is the 1-based step count."""
    p, g, m, v = (np.asarray(a, f32) for a in (p, g, m, v))
    m = (m + f32(1 - b1) * (g - m).astype(f32)).astype(f32)
    v = ((v * f32(b2)).astype(f32) + (f32(1 - b2) * g).astype(f32) * g).astype(f32)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = f32(lr / bc1)
    denom = ((np.sqrt(v).astype(f32) / f32(math.sqrt(bc2))).astype(f32) + f32(eps)).astype(f32)
    p = (p - step_size * (m / denom).astype(f32)).astype(f32)
    return p, m, v


def epoch_frame_parallel(beta, m, v, batches, grad_fn, first_step=1, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """Restatement of the schedule behind dnmf_motion_epoch's one-launch path: an epoch whose minibatches
    `batches` (lists of frame ids, no frame twice) touch disjoint columns of beta[10,3,T] is run column by
    column -- the zero-gradient Adam steps of the minibatches before the frame's own (the reference's
    dense Adam moves every column at every step, Demix/dNMF.py:190-191), its gradient step, the zero-gradient
    steps after.  `grad_fn(beta_columns[10,3,B], ids) -> grad[10,3,B]` stands for the fused kernel.  Must equal
    the step-by-step schedule (adam_dense over the whole tensor per minibatch) bit for bit."""
    beta, m, v = (np.array(a, f32) for a in (beta, m, v))
    n = len(batches)
    batch_of = {}
    for i, ids in enumerate(batches):
        for t in ids:
            assert t not in batch_of, "a frame drawn twice couples its steps: run batch by batch"
            batch_of[int(t)] = i
    kw = dict(lr=lr, b1=b1, b2=b2, eps=eps)
    zero = np.zeros((10, 3), f32)
    for t in range(beta.shape[2]):           # phase 0: replay the steps that precede the frame's minibatch
        for s_ in range(batch_of.get(t, n)):
            beta[:, :, t], m[:, :, t], v[:, :, t] = adam_dense(beta[:, :, t], zero, m[:, :, t], v[:, :, t],
                                                               first_step + s_, **kw)
    ids_all = [int(t) for ids in batches for t in ids]
    grads = np.concatenate([np.asarray(grad_fn(beta[:, :, list(ids)], list(ids)), f32) for ids in batches], axis=2)
    for j, t in enumerate(ids_all):          # phase 1: the gradient step, then the steps that follow
        g = grads[:, :, j]
        for s_ in range(batch_of[t], n):
            beta[:, :, t], m[:, :, t], v[:, :, t] = adam_dense(beta[:, :, t], g, m[:, :, t], v[:, :, t],
                                                               first_step + s_, **kw)
            g = zero
    return beta, m, v


def closed_form_mu_stats(frames: np.ndarray, times, beta, tabs, sz):
    """G[T',K,K], b[T',K] in float64 from the closed-form footprints (A_t values rounded to
    fp32 first, accumulated in fp64 like Demix/dNMF.py:141-142)."""
    K = tabs[0].shape[0]
    Gm = np.zeros((len(times), K, K))
    bv = np.zeros((len(times), K))
    for j, t in enumerate(times):
        ix = sample_coords(beta[:, :, t], sz)
        a = []
        for d in range(3):
            s = int(sz[d])
            c = np.fmin(np.fmax(ix[d], f32(-2)), f32(s))
            fl = np.floor(c)
            e = tabs[d][:, fl.astype(np.int32) + 2]           # [K,X,Y,Z,2]
            a.append((e[..., 0] + (c - fl).astype(f32)[None] * e[..., 1]).astype(f32))
        At = ((a[0] * a[1]).astype(f32) * a[2]).astype(f32).reshape(K, -1).astype(np.float64)
        Gm[j] = At @ At.T
        bv[j] = At @ frames[j].reshape(-1).astype(np.float64)
    return Gm, bv


def _cells(beta_t, sz, q_order="sequential"):
    """Cell index (+2: table row) and fraction of every voxel's sample, per axis, as closed_form_frame does."""
    ix = sample_coords(beta_t, sz, q_order)
    i_idx, fr = [], []
    for d in range(3):
        s = int(sz[d])
        c = np.fmin(np.fmax(ix[d], f32(-2)), f32(s))
        fl = np.floor(c)
        i_idx.append(fl.astype(np.int32) + 2)
        fr.append((c - fl).astype(f32))
    return i_idx, fr


def neuron_boxes(i_idx, rng):
    """Voxel boxes [K,3,2] (inclusive, hi < lo when empty) outside which a neuron's resampled footprint is exactly
    zero for this frame: table row i (= cell i - 2) is non-zero only for lo_k - 1 <= cell <= hi_k, so on axis d the
    voxel coordinate is kept when some voxel of that coordinate plane has its cell in that interval.  A superset of
    the support, which makes the boxed closed forms below bit-identical to the all-voxel ones (the skipped terms
    are +0)."""
    K = rng.shape[0]
    out = np.zeros((K, 3, 2), np.int64)
    for d in range(3):
        other = tuple(a for a in range(3) if a != d)
        cmin = i_idx[d].min(axis=other).astype(np.int64) - 2
        cmax = i_idx[d].max(axis=other).astype(np.int64) - 2
        for k in range(K):
            lo, hi = int(rng[k, d, 0]), int(rng[k, d, 1])
            live = np.nonzero((cmax >= lo - 1) & (cmin <= hi))[0] if lo <= hi else np.zeros(0, np.int64)
            out[k, d] = (live[0], live[-1]) if live.size else (0, -1)
    return out


def _footprint_in_box(k, box, i_idx, fr, tabs):
    """(a0, a1, a2, d0, d1, d2) of neuron k on its voxel box: per-axis lerped values and table differences."""
    sl = tuple(slice(int(box[d, 0]), int(box[d, 1]) + 1) for d in range(3))
    a, dd = [], []
    for d in range(3):
        e = tabs[d][k][i_idx[d][sl]]
        a.append((e[..., 0] + fr[d][sl] * e[..., 1]).astype(f32))
        dd.append(e[..., 1])
    return sl, a, dd


def closed_form_frame_boxed(frame, beta_t, c_t, tabs, rng, sz, q_order="sequential"):
    """closed_form_frame evaluated neuron by neuron on `neuron_boxes` only: same values bit for bit, but the cost
    is the number of in-support (voxel, neuron) pairs instead of N*K -- what makes the BASELINE configurations
    (cfg2 256x128x21 K=150, cfg3 512x256x32 K=300, cfg4 K=1000 sigma=6) checkable in seconds."""
    X, Y, Z = (int(s) for s in sz)
    i_idx, fr = _cells(beta_t, sz, q_order)
    boxes = neuron_boxes(i_idx, rng)
    yhat = np.zeros((X, Y, Z), f32)
    g = [np.zeros((X, Y, Z), f32) for _ in range(3)]
    for k in range(tabs[0].shape[0]):
        if boxes[k, :, 1].min() < 0 or (boxes[k, :, 1] < boxes[k, :, 0]).any():
            continue
        ck = f32(c_t[k])
        sl, a, dd = _footprint_in_box(k, boxes[k], i_idx, fr, tabs)
        yhat[sl] += ck * a[0] * a[1] * a[2]
        g[0][sl] += ck * dd[0] * a[1] * a[2]
        g[1][sl] += ck * a[0] * dd[1] * a[2]
        g[2][sl] += ck * a[0] * a[1] * dd[2]
    r = (yhat - frame.astype(f32)).astype(f32)
    sse = float(np.sum(r.astype(np.float64) ** 2))
    x = np.arange(X, dtype=np.float64)[:, None, None]
    y = np.arange(Y, dtype=np.float64)[None, :, None]
    z = np.arange(Z, dtype=np.float64)[None, None, :]
    grad = np.zeros((10, 3))
    for b in range(3):
        h = r.astype(np.float64) * g[b].astype(np.float64)
        hx, hy, hz = h.sum((1, 2)), h.sum((0, 2)), h.sum((0, 1))
        xs, ys, zs = x.ravel(), y.ravel(), z.ravel()
        grad[0, b] = 2.0 * h.sum()
        grad[1, b], grad[2, b], grad[3, b] = 2.0 * (xs @ hx), 2.0 * (ys @ hy), 2.0 * (zs @ hz)
        grad[4, b], grad[5, b], grad[6, b] = 2.0 * (xs * xs @ hx), 2.0 * (ys * ys @ hy), 2.0 * (zs * zs @ hz)
        grad[7, b] = 2.0 * (xs @ h.sum(2) @ ys)
        grad[8, b] = 2.0 * (xs @ h.sum(1) @ zs)
        grad[9, b] = 2.0 * (ys @ h.sum(0) @ zs)
    return yhat, sse, grad


def closed_form_step_boxed(frames, times, beta, C, tabs, rng, sz, q_order="sequential"):
    """closed_form_step on the neuron boxes (Demix/dNMF.py:187-190; mean over B*N)."""
    B = len(times)
    N = int(np.prod([int(s) for s in sz]))
    grad = np.zeros(beta.shape, np.float64)
    sse = 0.0
    for j, t in enumerate(times):
        _, s, g = closed_form_frame_boxed(frames[j], beta[:, :, t], C[:, t], tabs, rng, sz, q_order)
        sse += s
        grad[:, :, t] += g / (B * N)
    return sse / (B * N), grad.astype(f32)


def closed_form_mu_stats_boxed(frames, times, beta, tabs, rng, sz, slab: int = 16, q_order="sequential"):
    """closed_form_mu_stats (G_t = A_t^T A_t, b_t = A_t^T Y_t in fp64 of fp32 footprint values,
    Demix/dNMF.py:141-142) assembled x-slab by x-slab from the neuron boxes, with one BLAS product per slab over
    the neurons that reach it."""
    X, Y, Z = (int(s) for s in sz)
    K = tabs[0].shape[0]
    Gm = np.zeros((len(times), K, K))
    bv = np.zeros((len(times), K))
    for j, t in enumerate(times):
        i_idx, fr = _cells(beta[:, :, t], sz, q_order)
        boxes = neuron_boxes(i_idx, rng)
        live = [k for k in range(K) if (boxes[k, :, 1] >= boxes[k, :, 0]).all()]
        for x0 in range(0, X, slab):
            x1 = min(x0 + slab, X)
            act = [k for k in live if boxes[k, 0, 0] < x1 and boxes[k, 0, 1] >= x0]
            if not act:
                continue
            panel = np.zeros((len(act), x1 - x0, Y, Z))
            for r_, k in enumerate(act):
                box = boxes[k].copy()
                box[0, 0], box[0, 1] = max(box[0, 0], x0), min(box[0, 1], x1 - 1)
                sl, a, _ = _footprint_in_box(k, box, i_idx, fr, tabs)
                val = ((a[0] * a[1]).astype(f32) * a[2]).astype(f32)
                panel[(r_, slice(sl[0].start - x0, sl[0].stop - x0), sl[1], sl[2])] = val
            P = panel.reshape(len(act), -1)
            Gm[j][np.ix_(act, act)] += P @ P.T
            bv[j][act] += P @ frames[j][x0:x1].reshape(-1).astype(np.float64)
    return Gm, bv


# ------------------------------------------------------------------------------------------------
# 3. Binning spec (integer, bit-exact target for the binning kernel)
# ------------------------------------------------------------------------------------------------


def tile_grid(sz: Sequence[int], tile: Sequence[int]) -> Tuple[int, int, int]:
    return tuple((int(s) + int(t) - 1) // int(t) for s, t in zip(sz, tile))


def tile_window(beta_t: np.ndarray, box_lo: Sequence[int], box_hi: Sequence[int],
                sz: Sequence[int]) -> np.ndarray:
    """Conservative window [3,2] = (wlo, whi) of table-entry indices i that voxels of the
    inclusive box can touch, by fp32 interval arithmetic over the 10 monomials in basis order.

    lo = b0; hi = b0; for a = 1..9: p1 = fl(b_a*mlo_a), p2 = fl(b_a*mhi_a);
    lo = fl(lo + min(p1,p2)); hi = fl(hi + max(p1,p2)).
    wlo = floor(clamp(lo, -4, s+4)) - 1, whi = floor(clamp(hi, -4, s+4)) + 1, both clamped
    to the table domain [-2, s]."""
    b = np.asarray(beta_t, f32)
    x0, y0, z0 = (f32(v) for v in box_lo)
    x1, y1, z1 = (f32(v) for v in box_hi)
    mlo = [f32(1), x0, y0, z0, f32(x0 * x0), f32(y0 * y0), f32(z0 * z0), f32(x0 * y0), f32(x0 * z0), f32(y0 * z0)]
    mhi = [f32(1), x1, y1, z1, f32(x1 * x1), f32(y1 * y1), f32(z1 * z1), f32(x1 * y1), f32(x1 * z1), f32(y1 * z1)]
    win = np.zeros((3, 2), np.int32)
    with np.errstate(all="ignore"):
        for d in range(3):
            lo = f32(b[0, d])
            hi = f32(b[0, d])
            for a in range(1, 10):
                p1 = f32(b[a, d] * mlo[a])
                p2 = f32(b[a, d] * mhi[a])
                lo = f32(lo + np.fmin(p1, p2))
                hi = f32(hi + np.fmax(p1, p2))
            s = int(sz[d])
            lo = np.fmin(np.fmax(lo, f32(-4)), f32(s + 4))
            hi = np.fmin(np.fmax(hi, f32(-4)), f32(s + 4))
            wlo = int(np.floor(lo)) - 1
            whi = int(np.floor(hi)) + 1
            win[d, 0] = min(max(wlo, -2), s)
            win[d, 1] = min(max(whi, -2), s)
    return win


def bin_tiles(beta: np.ndarray, times: Sequence[int], rng: np.ndarray, sz, tile):
    """Per (batch slot, tile) sorted neuron lists.  Returns (counts[B*nt], offsets[B*nt+1],
    ids[total], windows[B*nt,3,2]).  Tile id = (bz*nty + by)*ntx + bx.  Neuron k is listed iff
    its range is non-empty and, on every axis, lo_k <= whi+1 and hi_k >= wlo."""
    ntx, nty, ntz = tile_grid(sz, tile)
    nt = ntx * nty * ntz
    counts, ids, wins = [], [], []
    for t in times:
        for bz in range(ntz):
            for by in range(nty):
                for bx in range(ntx):
                    lo = (bx * tile[0], by * tile[1], bz * tile[2])
                    hi = tuple(min(lo[d] + tile[d], int(sz[d])) - 1 for d in range(3))
                    w = tile_window(beta[:, :, t], lo, hi, sz)
                    ok = np.ones(rng.shape[0], bool)
                    for d in range(3):
                        ok &= rng[:, d, 0] <= rng[:, d, 1]
                        ok &= rng[:, d, 0] <= w[d, 1] + 1
                        ok &= rng[:, d, 1] >= w[d, 0]
                    k = np.nonzero(ok)[0].astype(np.int32)
                    counts.append(len(k))
                    ids.append(k)
                    wins.append(w)
    counts = np.asarray(counts, np.int32)
    offsets = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    ids = np.concatenate(ids) if ids else np.zeros(0, np.int32)
    return counts, offsets, ids, np.asarray(wins, np.int32).reshape(len(times) * nt, 3, 2)


# ------------------------------------------------------------------------------------------------
# 4. EXTENSION oracle (no reference counterpart): learnable positions, widths and scalar background
# ------------------------------------------------------------------------------------------------


class ExtendedPort:
    """torch-autograd restatement of the extended model Yhat = sum_k C_k A_t(pos, sigma) + b, where the
    discretised Gaussian volume of Demix/dNMF.py:39-40 is rebuilt from LEAF pos / sigma on every forward
    (the reference keeps them fixed).  Used only to test the shared-parameter gradient kernel; every test
    built on it is labelled "extension, not reference parity"."""

    def __init__(self, sz, positions, sigma, C, beta, background=0.0):
        self.sz = torch.as_tensor(sz).long()
        self.grid_id, self.phi = voxel_basis(self.sz.tolist())
        self.pos = torch.as_tensor(positions).float().clone().requires_grad_(True)
        self.sigma = torch.as_tensor(sigma).float().clone().requires_grad_(True)
        self.bg = torch.tensor(float(background), requires_grad=True)
        self.beta = torch.as_tensor(beta).float().clone().requires_grad_(True)
        self.C = torch.as_tensor(C).float()

    def loss(self, frames: torch.Tensor, times: Sequence[int]) -> torch.Tensor:
        times = list(times)
        A = gaussian_volume(self.grid_id, self.pos, self.sigma)
        q = torch.einsum("mnza,abt->mnzbt", self.phi, self.beta[:, :, times])
        u = 2 * q / (self.sz[None, None, None, :, None] - 1) - 1
        vol = A.permute(3, 2, 1, 0)[None].expand(len(times), -1, -1, -1, -1)
        A_t = F.grid_sample(vol, u.permute(4, 2, 1, 0, 3), mode="bilinear", padding_mode="zeros",
                            align_corners=True).permute(0, 1, 4, 3, 2)
        yhat = torch.einsum("tkmnz,kt->tmnz", A_t, self.C[:, times]) + self.bg
        return F.mse_loss(yhat, frames)

    def grads(self, frames, times):
        for t in (self.pos, self.sigma, self.bg, self.beta):
            t.grad = None
        loss = self.loss(frames, times)
        loss.backward()
        return (float(loss.detach()), self.beta.grad.numpy(), self.pos.grad.numpy(), self.sigma.grad.numpy(),
                float(self.bg.grad))
