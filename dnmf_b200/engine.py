"""Thin torch-tensor front end of the C ABI (include/dnmf_b200.h).

torch is used only for device memory, streams and tensor hand-off: every method turns tensors
into raw pointers and calls the CUDA library through ctypes.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_dev(t: torch.Tensor, dtype, name: str, device=None, shape=None):
    if not torch.is_tensor(t) or not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise _lib.DnmfError("%s must be a contiguous CUDA %s tensor" % (name, dtype))
    if device is not None and t.device != device:
        raise _lib.DnmfError("%s is on %s, engine is on %s" % (name, t.device, device))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise _lib.DnmfError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))


class Engine:
    """One GPU's slab: T local frames of an X*Y*Z volume with K neurons."""

    def __init__(self, sz: Sequence[int], K: int, T: int, device=None):
        if not torch.cuda.is_available():
            raise _lib.DnmfError("dnmf_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.X, self.Y, self.Z = (int(s) for s in sz)
        self.K, self.T = int(K), int(T)
        self.N = self.X * self.Y * self.Z
        h = ctypes.c_void_p()
        _lib.check(self.lib.dnmf_create(ctypes.byref(h), self.X, self.Y, self.Z, self.K, self.T,
                                        self.device.index), "dnmf_create")
        self._h = h
        self._has_video = False
        self._attached = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dnmf_destroy(self._h)
            self._h = None
        self._attached = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return _stream_ptr(self.device)

    def _ids32(self, frame_ids, allow_duplicates: bool = True) -> torch.Tensor:
        """int32 CUDA ids of this engine.  Ids that are still on the host are validated here (range, duplicates)
        so that the error is immediate; ids already on the device are checked by the library's own kernel."""
        t = torch.as_tensor(frame_ids)
        if not t.is_cuda:
            flat = t.reshape(-1).to(torch.int64)
            if flat.numel() == 0:
                raise _lib.DnmfError("empty frame id batch")
            lo, hi = int(flat.min()), int(flat.max())
            if lo < 0 or hi >= self.T:
                raise _lib.DnmfError("frame id %d outside [0, %d) (ids index this engine's slab of T frames)"
                                     % (lo if lo < 0 else hi, self.T))
            if not allow_duplicates and int(torch.unique(flat).numel()) != int(flat.numel()):
                raise _lib.DnmfError("a frame id occurs twice in the batch")
        return t.reshape(-1).to(self.device, torch.int32).contiguous()

    def _check_state(self, beta, m=None, v=None, C=None):
        """Shape / dtype / device of the tensors whose raw pointers go to the library."""
        _check_dev(beta, torch.float32, "beta", self.device, (10, 3, self.T))
        if m is not None:
            _check_dev(m, torch.float32, "exp_avg", self.device, (10, 3, self.T))
        if v is not None:
            _check_dev(v, torch.float32, "exp_avg_sq", self.device, (10, 3, self.T))
        if C is not None:
            _check_dev(C, torch.float32, "C", self.device, (self.K, self.T))

    def check_status(self):
        """Synchronises the stream and raises if an earlier asynchronous call found an error on the device."""
        _lib.check(self.lib.dnmf_check_status(self._h, self.stream), "dnmf_check_status")

    # -- footprints / tiling ----------------------------------------------------------------------
    def set_footprints(self, pos, sigma, cutoff: float):
        pos = np.ascontiguousarray(torch.as_tensor(pos).detach().cpu().numpy(), np.float32)
        sigma = np.ascontiguousarray(torch.as_tensor(sigma).detach().cpu().numpy(), np.float32)
        if pos.shape != (self.K, 3) or sigma.shape != (self.K,):
            raise _lib.DnmfError("positions must be [K,3] and sigma [K]")
        _lib.check(self.lib.dnmf_set_footprints(self._h, pos.ctypes.data_as(ctypes.c_void_p),
                                                sigma.ctypes.data_as(ctypes.c_void_p), float(cutoff), self.stream),
                   "dnmf_set_footprints")

    def ranges(self) -> np.ndarray:
        out = np.zeros((self.K, 3, 2), np.int32)
        _lib.check(self.lib.dnmf_get_ranges(self._h, out.ctypes.data_as(ctypes.c_void_p)), "dnmf_get_ranges")
        return out

    def table(self, axis: int) -> np.ndarray:
        s = (self.X, self.Y, self.Z)[axis]
        out = np.zeros((self.K, s + 3, 2), np.float32)
        _lib.check(self.lib.dnmf_get_table(self._h, axis, out.ctypes.data_as(ctypes.c_void_p)), "dnmf_get_table")
        return out

    def set_tiling(self, warps_x: int = 1, warps_y: int = 1, tz: int = 0, slot_capacity: int = 0, subtiles_y: int = 1,
                   warps_z: int = 1):
        _lib.check(self.lib.dnmf_set_tiling(self._h, warps_x, warps_y, tz, slot_capacity, subtiles_y, warps_z),
                   "dnmf_set_tiling")

    def set_affine(self, affine: bool):
        """Affine fit: loss_grad leaves the (frozen) quadratic gradient rows zero and affine frames take the shorter
        main loop (dnmf_set_affine)."""
        _lib.check(self.lib.dnmf_set_affine(self._h, int(bool(affine))), "dnmf_set_affine")

    def tiling(self) -> dict:
        out = np.zeros(12, np.int32)
        _lib.check(self.lib.dnmf_get_tiling(self._h, out.ctypes.data_as(ctypes.c_void_p)), "dnmf_get_tiling")
        keys = ("tx", "ty", "tz", "ntx", "nty", "ntz", "warps_x", "warps_y", "cap", "subtiles_y", "fast_div", "warps_z")
        return dict(zip(keys, (int(v) for v in out)))

    # -- video ------------------------------------------------------------------------------------
    def upload_frames(self, frames: torch.Tensor, t0: int = 0, clamp_negative: bool = True):
        """frames: host float32 [n,X,Y,Z] (pinned memory makes the copy asynchronous)."""
        if frames.is_cuda:
            vid = self.video()
            vid[t0:t0 + frames.shape[0]].copy_(frames.clamp(min=0) if clamp_negative else frames)
        else:
            frames = frames.contiguous().float()
            _lib.check(self.lib.dnmf_upload_frames(self._h, _ptr(frames), int(t0), int(frames.shape[0]),
                                                   int(clamp_negative), self.stream), "dnmf_upload_frames")
            torch.cuda.current_stream(self.device).synchronize()
            self._attached = None      # the library switched to a slab of its own
        self._has_video = True

    def attach_frames(self, frames: torch.Tensor, clamp_negative: bool = True):
        """Zero-copy: the engine reads the caller's CUDA slab [T,X,Y,Z] in place (clamped in place when asked, like
        the reference's dataset does to its own video, Demix/dNMF.py:215).  The engine keeps a reference."""
        _check_dev(frames, torch.float32, "frames", self.device, (self.T, self.X, self.Y, self.Z))
        _lib.check(self.lib.dnmf_attach_frames(self._h, _ptr(frames), int(clamp_negative), self.stream),
                   "dnmf_attach_frames")
        self._attached = frames
        self._has_video = True

    def video(self) -> torch.Tensor:
        """Zero-copy torch view [T,X,Y,Z] of the resident slab."""
        p = ctypes.c_void_p()
        _lib.check(self.lib.dnmf_video_devptr(self._h, ctypes.byref(p)), "dnmf_video_devptr")

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.T, self.X, self.Y, self.Z), "typestr": "<f4",
                                      "data": (p.value, False), "version": 2}
        self._has_video = True
        return torch.as_tensor(a, device=self.device)

    # -- kernels ----------------------------------------------------------------------------------
    def bin_tiles(self, beta: torch.Tensor, frame_ids: torch.Tensor):
        """Stand-alone binning.  Returns (counts, offsets, ids, windows) as numpy arrays."""
        self._check_state(beta)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        tl = self.tiling()
        nt = tl["ntx"] * tl["nty"] * tl["ntz"]
        counts = torch.zeros(B * nt, dtype=torch.int32, device=self.device)
        offsets = torch.zeros(B * nt + 1, dtype=torch.int64, device=self.device)
        windows = torch.zeros(B * nt, 3, 2, dtype=torch.int32, device=self.device)
        total = ctypes.c_int64(0)
        _lib.check(self.lib.dnmf_bin_tiles(self._h, _ptr(beta), _ptr(ids32), B, _ptr(counts), _ptr(offsets),
                                           _ptr(windows), None, 0, ctypes.byref(total), self.stream), "dnmf_bin_tiles")
        ids = torch.zeros(max(1, total.value), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.dnmf_bin_tiles(self._h, _ptr(beta), _ptr(ids32), B, _ptr(counts), _ptr(offsets),
                                           _ptr(windows), _ptr(ids), int(ids.numel()), ctypes.byref(total),
                                           self.stream), "dnmf_bin_tiles")
        return (counts.cpu().numpy(), offsets.cpu().numpy(), ids[:total.value].cpu().numpy(), windows.cpu().numpy())

    def loss_grad(self, frame_ids: torch.Tensor, beta: torch.Tensor, C: torch.Tensor,
                  frames: Optional[torch.Tensor] = None, B_global: Optional[int] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (grad[10,3,T] float32 with only the batch columns non-zero, sse[B] float64)."""
        self._check_state(beta, C=C)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (B, self.X, self.Y, self.Z))
        grad = torch.zeros(10, 3, self.T, dtype=torch.float32, device=self.device)
        sse = torch.zeros(B, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.dnmf_loss_grad(self._h, _ptr(frames), _ptr(ids32), B, int(B_global or B), _ptr(beta),
                                           _ptr(C), _ptr(grad), _ptr(sse), self.stream), "dnmf_loss_grad")
        return grad, sse

    def adam_step(self, beta, grad, m, v, lr, betas, eps, step, affine=False):
        for t, n in ((beta, "beta"), (grad, "grad"), (m, "exp_avg"), (v, "exp_avg_sq")):
            _check_dev(t, torch.float32, n)
        _lib.check(self.lib.dnmf_adam_step(self._h, _ptr(beta), _ptr(grad), _ptr(m), _ptr(v), float(lr),
                                           float(betas[0]), float(betas[1]), float(eps), int(step), int(affine),
                                           None, 0, 0, None, self.stream), "dnmf_adam_step")

    def motion_step(self, frame_ids32: torch.Tensor, beta, m, v, C, lr, betas, eps, step, affine=False,
                    frames: Optional[torch.Tensor] = None, B_global: Optional[int] = None,
                    loss_out: Optional[torch.Tensor] = None):
        """Device-resident step (frames=None reads the resident slab).  loss_out: float64 CUDA scalar."""
        self._check_state(beta, m, v, C)
        frame_ids32 = self._ids32(frame_ids32)
        B = int(frame_ids32.numel())
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (B, self.X, self.Y, self.Z))
        if loss_out is not None:
            _check_dev(loss_out, torch.float64, "loss_out", self.device)
        _lib.check(self.lib.dnmf_motion_step(self._h, _ptr(frames), _ptr(frame_ids32), B, int(B_global or B),
                                             _ptr(beta), _ptr(m), _ptr(v), _ptr(C), float(lr), float(betas[0]),
                                             float(betas[1]), float(eps), int(step), int(affine), _ptr(loss_out),
                                             self.stream), "dnmf_motion_step")

    def motion_epoch(self, ids_dev: torch.Tensor, offsets, beta, m, v, C, lr, betas, eps, first_step, affine=False,
                     global_batch_scale: int = 1, loss_out: Optional[torch.Tensor] = None):
        """All minibatches of an epoch over the resident video in one library call.  ids_dev: int32 CUDA tensor with
        the batches concatenated; offsets: nbatches+1 host ints; loss_out: float64 CUDA tensor [nbatches]."""
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32))
        nb = int(off.size) - 1
        self._check_state(beta, m, v, C)
        _check_dev(ids_dev, torch.int32, "ids_dev", self.device)
        if nb < 0 or int(off[0]) < 0 or int(off[-1]) > int(ids_dev.numel()) or np.any(np.diff(off) < 1):
            raise _lib.DnmfError("batch offsets must be increasing and stay inside the id array")
        if loss_out is not None:
            _check_dev(loss_out, torch.float64, "loss_out", self.device, (nb,))
        _lib.check(self.lib.dnmf_motion_epoch(self._h, _ptr(ids_dev), ctypes.c_void_p(off.ctypes.data), nb,
                                              int(global_batch_scale), _ptr(beta), _ptr(m), _ptr(v), _ptr(C),
                                              float(lr), float(betas[0]), float(betas[1]), float(eps),
                                              int(first_step), int(affine), _ptr(loss_out), self.stream),
                   "dnmf_motion_epoch")

    def epoch_mode(self, sequential: int = -1) -> int:
        """Select (1 = batch by batch, 0 = automatic) and/or query how `motion_epoch` runs: returns 1 when the last
        call ran frame-parallel (one fused launch over all frames of the epoch)."""
        last = ctypes.c_int(0)
        _lib.check(self.lib.dnmf_epoch_mode(self._h, int(sequential), ctypes.byref(last)), "dnmf_epoch_mode")
        return int(last.value)

    def motion_step_host(self, frames_host: torch.Tensor, ids_host: torch.Tensor, beta, m, v, C, lr, betas, eps,
                         step, affine=False, B_global: Optional[int] = None) -> float:
        """End-to-end step from HOST buffers (H2D copy + kernels + loss read-back)."""
        self._check_state(beta, m, v, C)
        B = int(ids_host.numel())
        if frames_host.is_cuda or frames_host.dtype != torch.float32 or not frames_host.is_contiguous() or \
                tuple(frames_host.shape) != (B, self.X, self.Y, self.Z):
            raise _lib.DnmfError("frames_host must be a contiguous float32 host tensor [B,X,Y,Z]")
        ids_host = torch.as_tensor(ids_host).reshape(-1).to("cpu", torch.int32).contiguous()
        loss = ctypes.c_double(0.0)
        _lib.check(self.lib.dnmf_motion_step_host(self._h, _ptr(frames_host), _ptr(ids_host), B, int(B_global or B),
                                                  _ptr(beta), _ptr(m), _ptr(v), _ptr(C), float(lr), float(betas[0]),
                                                  float(betas[1]), float(eps), int(step), int(affine),
                                                  ctypes.byref(loss), self.stream), "dnmf_motion_step_host")
        return loss.value

    def forward(self, frame_ids: torch.Tensor, beta, C, want_At=False, want_grid=False):
        self._check_state(beta, C=C)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        AtC = torch.empty(B, self.X, self.Y, self.Z, dtype=torch.float32, device=self.device)
        At = torch.empty(B, self.K, self.X, self.Y, self.Z, dtype=torch.float32, device=self.device) if want_At else None
        grid = torch.empty(self.X, self.Y, self.Z, 3, B, dtype=torch.float32, device=self.device) if want_grid else None
        _lib.check(self.lib.dnmf_forward(self._h, _ptr(ids32), B, _ptr(beta), _ptr(C), _ptr(AtC), _ptr(At),
                                         _ptr(grid), self.stream), "dnmf_forward")
        return AtC, At, grid

    def mu_stats(self, frame_ids: torch.Tensor, beta, frames: Optional[torch.Tensor] = None):
        self._check_state(beta)
        ids32 = self._ids32(frame_ids, allow_duplicates=False)
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (int(ids32.numel()), self.X, self.Y, self.Z))
        _lib.check(self.lib.dnmf_mu_stats(self._h, _ptr(frames), _ptr(ids32), int(ids32.numel()), _ptr(beta),
                                          self.stream), "dnmf_mu_stats")

    def mu_path(self, flags: int = -1) -> int:
        """Select (bit 0 = skip the fused tiles, bit 1 = dense sweeps only, bit 2 = one launch per sweep, bit 3 =
        tensor-core panel kernel first, 0 = automatic) and/or query the device paths of the trace update: returns
        bit 0 = the last `mu_stats` ran on the fused kernel's tiles, bit 1 = the last `mu_begin` / `mu_sweeps` used
        the neighbour-compacted statistics, bit 2 = the last `mu_stats` ran on the tensor-core panel kernel."""
        last = ctypes.c_int(0)
        _lib.check(self.lib.dnmf_mu_path(self._h, int(flags), ctypes.byref(last)), "dnmf_mu_path")
        return int(last.value)

    def get_mu_stats(self, t: int):
        G = np.zeros((self.K, self.K))
        b = np.zeros(self.K)
        _lib.check(self.lib.dnmf_get_mu_stats(self._h, int(t), G.ctypes.data_as(ctypes.c_void_p),
                                              b.ctypes.data_as(ctypes.c_void_p)), "dnmf_get_mu_stats")
        return G, b

    def mu_sweeps(self, C: torch.Tensor, gamma, iters: int):
        _check_dev(C, torch.float32, "C", self.device, (self.K, self.T))
        _lib.check(self.lib.dnmf_mu_sweeps(self._h, _ptr(C), float(gamma or 0.0), int(gamma is not None), int(iters),
                                           self.stream), "dnmf_mu_sweeps")

    def mu_begin(self, C):
        _check_dev(C, torch.float32, "C", self.device, (self.K, self.T))
        _lib.check(self.lib.dnmf_mu_begin(self._h, _ptr(C), self.stream), "dnmf_mu_begin")

    def mu_sweep(self, gamma, halo_prev=None, halo_next=None):
        _lib.check(self.lib.dnmf_mu_sweep(self._h, float(gamma or 0.0), int(gamma is not None), _ptr(halo_prev),
                                          _ptr(halo_next), self.stream), "dnmf_mu_sweep")

    def mu_boundary(self):
        first = torch.empty(self.K, dtype=torch.float64, device=self.device)
        last = torch.empty(self.K, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.dnmf_mu_boundary(self._h, _ptr(first), _ptr(last), self.stream), "dnmf_mu_boundary")
        return first, last

    def mu_end(self, C):
        _check_dev(C, torch.float32, "C", self.device, (self.K, self.T))
        _lib.check(self.lib.dnmf_mu_end(self._h, _ptr(C), self.stream), "dnmf_mu_end")

    def iwarp(self, frame_ids: torch.Tensor, beta, frames: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check_state(beta)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (B, self.X, self.Y, self.Z))
        out = torch.empty(B, self.X, self.Y, self.Z, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dnmf_iwarp(self._h, _ptr(frames), _ptr(ids32), B, _ptr(beta), _ptr(out), self.stream),
                   "dnmf_iwarp")
        return out

    def forward_maxz(self, frame_ids: torch.Tensor, beta) -> torch.Tensor:
        """max over z of the deformed footprints, [B,K,X,Y] (demo.py:50-52 `A_t.max(2)`), without the dense A_t."""
        self._check_state(beta)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        out = torch.empty(B, self.K, self.X, self.Y, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dnmf_forward_maxz(self._h, _ptr(ids32), B, _ptr(beta), _ptr(out), self.stream),
                   "dnmf_forward_maxz")
        return out

    def frames_maxz(self, frames: torch.Tensor) -> torch.Tensor:
        """max over the last axis of a contiguous float32 CUDA tensor [..., Z] (Y.max(2), Y_i.max(2))."""
        _check_dev(frames, torch.float32, "frames", self.device)
        out = torch.empty(frames.shape[:-1], dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dnmf_frames_maxz(_ptr(frames), int(out.numel()), int(frames.shape[-1]), _ptr(out),
                                             self.stream), "dnmf_frames_maxz")
        return out

    # -- extension: shared-parameter gradients (no reference counterpart) -----------------------------
    def ext_enable(self):
        _lib.check(self.lib.dnmf_ext_enable(self._h), "dnmf_ext_enable")

    def ext_loss_grad(self, frame_ids: torch.Tensor, beta, C, background: float = 0.0,
                      frames: Optional[torch.Tensor] = None, B_global: Optional[int] = None, grad_beta=None):
        """Returns (grad_beta[10,3,T], sse[B], gpos[K,3], gsig[K], gbg[1]) -- the last three in float64."""
        self._check_state(beta, C=C)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (B, self.X, self.Y, self.Z))
        if grad_beta is None:
            grad_beta = torch.zeros(10, 3, self.T, dtype=torch.float32, device=self.device)
        sse = torch.zeros(B, dtype=torch.float64, device=self.device)
        gpos = torch.zeros(self.K, 3, dtype=torch.float64, device=self.device)
        gsig = torch.zeros(self.K, dtype=torch.float64, device=self.device)
        gbg = torch.zeros(1, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.dnmf_ext_loss_grad(self._h, _ptr(frames), _ptr(ids32), B, int(B_global or B), _ptr(beta),
                                               _ptr(C), float(background), _ptr(grad_beta), _ptr(sse), _ptr(gpos),
                                               _ptr(gsig), _ptr(gbg), self.stream), "dnmf_ext_loss_grad")
        return grad_beta, sse, gpos, gsig, gbg

    # device-resident iteration of the extension (no host round trip; the caller all-reduces `packed` in between)
    def ext_step_begin(self, frame_ids: torch.Tensor, beta, C, packed: torch.Tensor,
                       frames: Optional[torch.Tensor] = None, B_global: Optional[int] = None):
        """packed: float64 CUDA tensor of 4K+2 entries <- (dL/dpos[K,3], dL/dsigma[K], dL/db, sum of the batch SSE)."""
        self._check_state(beta, C=C)
        ids32 = self._ids32(frame_ids)
        B = int(ids32.numel())
        if frames is not None:
            _check_dev(frames, torch.float32, "frames", self.device, (B, self.X, self.Y, self.Z))
        _check_dev(packed, torch.float64, "packed", self.device, (4 * self.K + 2,))
        _lib.check(self.lib.dnmf_ext_step_begin(self._h, _ptr(frames), _ptr(ids32), B, int(B_global or B), _ptr(beta),
                                                _ptr(C), _ptr(packed), self.stream), "dnmf_ext_step_begin")

    def ext_step_end(self, beta, m, v, lr, betas, eps, step, affine, packed: torch.Tensor, B_global: int,
                     lr_pos: float, lr_sigma: float, lr_background: float, sigma_min: float = 0.5,
                     loss_out: Optional[torch.Tensor] = None):
        self._check_state(beta, m, v)
        _check_dev(packed, torch.float64, "packed", self.device, (4 * self.K + 2,))
        if loss_out is not None:
            _check_dev(loss_out, torch.float64, "loss_out", self.device)
        _lib.check(self.lib.dnmf_ext_step_end(self._h, _ptr(beta), _ptr(m), _ptr(v), float(lr), float(betas[0]),
                                              float(betas[1]), float(eps), int(step), int(affine), _ptr(packed),
                                              int(B_global), float(lr_pos), float(lr_sigma), float(lr_background),
                                              float(sigma_min), _ptr(loss_out), self.stream), "dnmf_ext_step_end")

    def ext_set_params(self, background: float = 0.0, reset_adam_state: bool = False):
        _lib.check(self.lib.dnmf_ext_set_params(self._h, float(background), int(reset_adam_state), self.stream),
                   "dnmf_ext_set_params")

    def ext_get_params(self):
        """(pos[K,3], sigma[K], background[]) as the context holds them now (float32 CUDA tensors, copies)."""
        pos = torch.empty(self.K, 3, dtype=torch.float32, device=self.device)
        sigma = torch.empty(self.K, dtype=torch.float32, device=self.device)
        bg = torch.empty((), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dnmf_ext_get_params(self._h, _ptr(pos), _ptr(sigma), _ptr(bg), self.stream),
                   "dnmf_ext_get_params")
        return pos, sigma, bg

    def counters(self) -> dict:
        out = np.zeros(8, np.int64)
        _lib.check(self.lib.dnmf_get_counters(self._h, out.ctypes.data_as(ctypes.c_void_p)), "dnmf_get_counters")
        keys = ("fit_launches", "reduce_launches", "prepass_launches", "table_builds", "adam_launches",
                "dense_forward_launches", "mu_stats_launches", "mu_sweep_launches")
        return dict(zip(keys, (int(v) for v in out)))


# -- context-free entry points (dense fp64 multiplicative updates, synthetic generator) ---------------------
def _cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.DnmfError("dnmf_b200 needs a CUDA device: there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.DnmfError("device must be a CUDA device, got %s" % dev)
    return torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev


def update_spatial_dense(A, C, Y_i, D=None, gamma=None, positions=None, grid=None, device=None) -> torch.Tensor:
    """A <- A * (Y_i C^T) / (A (C C^T) + gamma D + 1e-32) in fp64 on the GPU (Demix/dNMF.py:151-160).  A[..., K],
    C[K, T], Y_i[..., T] with the same leading (voxel) axes.  D: array like A, or None; with `positions` [K,3] and
    `grid` (X, Y, Z) the penalty D = 1 - exp(-0.01 |p - pos_k|) (:133-135) is computed on the fly instead."""
    dev = _cuda_device(device)
    lib = _lib.load()
    Ad = torch.as_tensor(A, dtype=torch.float64).to(dev).contiguous()
    Cd = torch.as_tensor(C, dtype=torch.float64).to(dev).contiguous()
    Yd = torch.as_tensor(Y_i, dtype=torch.float64).to(dev).contiguous()
    K, T = int(Cd.shape[0]), int(Cd.shape[1])
    if Ad.shape[-1] != K or Yd.shape[-1] != T or Ad.shape[:-1] != Yd.shape[:-1]:
        raise _lib.DnmfError("update_spatial: A[...,K], C[K,T], Y_i[...,T] do not fit together")
    P = int(Ad.numel() // K)
    use_D, Dd, pos, g = 0, None, None, (0, 0, 0)
    if D is not None:
        if gamma is None:
            raise _lib.DnmfError("update_spatial: D needs gamma")   # the reference raises a TypeError here
        Dd = torch.as_tensor(D, dtype=torch.float64).to(dev).contiguous()
        if Dd.numel() != Ad.numel():
            raise _lib.DnmfError("update_spatial: D must have the shape of A")
        use_D = 1
    elif positions is not None and gamma is not None:
        g = tuple(int(v) for v in grid)
        pos = torch.as_tensor(positions, dtype=torch.float32).to(dev).contiguous()
        use_D = 2
    scratch = torch.empty(K * K, dtype=torch.float64, device=dev)
    out = torch.empty_like(Ad)
    with torch.cuda.device(dev):
        _lib.check(lib.dnmf_update_spatial(_ptr(Ad), _ptr(Cd), _ptr(Yd), _ptr(Dd), _ptr(pos), g[0], g[1], g[2],
                                           float(gamma or 0.0), use_D, P, K, T, _ptr(scratch), _ptr(out),
                                           _stream_ptr(dev)), "dnmf_update_spatial")
    return out


def update_temporal_dense(A_t, C, Y, gamma=None, device=None) -> torch.Tensor:
    """The static update_temporal on dense arrays (Demix/dNMF.py:139-149), fp64 on the GPU: A_t[..., K, T], C[K,T],
    Y[..., T]."""
    dev = _cuda_device(device)
    lib = _lib.load()
    Ad = torch.as_tensor(A_t, dtype=torch.float64).to(dev).contiguous()
    Cd = torch.as_tensor(C, dtype=torch.float64).to(dev).contiguous()
    Yd = torch.as_tensor(Y, dtype=torch.float64).to(dev).contiguous()
    K, T = int(Cd.shape[0]), int(Cd.shape[1])
    if tuple(Ad.shape[-2:]) != (K, T) or Yd.shape[-1] != T or Ad.shape[:-2] != Yd.shape[:-1]:
        raise _lib.DnmfError("update_temporal: A_t[...,K,T], C[K,T], Y[...,T] do not fit together")
    P = int(Yd.numel() // T)
    scratch = torch.empty((K * K + K) * T, dtype=torch.float64, device=dev)
    out = torch.empty_like(Cd)
    with torch.cuda.device(dev):
        _lib.check(lib.dnmf_update_temporal_dense(_ptr(Ad), _ptr(Cd), _ptr(Yd), float(gamma or 0.0),
                                                  int(gamma is not None), P, K, T, _ptr(scratch), _ptr(out),
                                                  _stream_ptr(dev)), "dnmf_update_temporal_dense")
    return out


def render_cells(positions, traces, sz, shape_std, device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Clean synthetic frames [T,X,Y,Z] = sum_k traces[k,t] exp(-|p - P_k(t)|^2 / (2 shape_std))
    (WUtils/Simulator.py:66-77,197-203) with the hand-written generator kernel."""
    dev = _cuda_device(device)
    lib = _lib.load()
    pos = torch.as_tensor(positions, dtype=torch.float32).to(dev).contiguous()
    tr = torch.as_tensor(np.asarray(traces), dtype=torch.float32).to(dev).contiguous()
    K, _, T = pos.shape
    X, Y, Z = (int(v) for v in sz)
    if out is None:
        out = torch.empty(T, X, Y, Z, dtype=torch.float32, device=dev)
    _check_dev(out, torch.float32, "out", dev, (T, X, Y, Z))
    with torch.cuda.device(dev):
        _lib.check(lib.dnmf_render_cells(_ptr(pos), _ptr(tr), int(K), int(T), 0, int(T), X, Y, Z, float(shape_std),
                                         _ptr(out), _stream_ptr(dev)), "dnmf_render_cells")
    return out
