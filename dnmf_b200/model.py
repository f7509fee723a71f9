"""Host-side mirror of the reference's class surface (Demix/dNMF.py) on top of the CUDA library.

Same names, constructor arguments, entry points and returned values as the reference:

    ExponentialFP(sz, K, T, positions=None, shape_std=3)           Demix/dNMF.py:18-122
        .beta [10,3,T] leaf tensor, .A [X,Y,Z,K], .pos, .sigma, .sz, forward(times, C)
    DeformableNMF(sz, K, T, positions=None)                         Demix/dNMF.py:124-194
        .fp, .C [K,T], update_motion(dataloader, optimizer, gamma, epochs),
        update_footprints(testloader, batch_size, sz, gamma_c, gamma_a, iter_c),
        static update_temporal / update_spatial

All arithmetic of the hot path runs in the CUDA kernels behind include/dnmf_b200.h; torch is used
for device memory, streams and tensor hand-off.  Extra keyword-only arguments (cutoff, deformation,
tiling, frame sharding) default to the reference's behaviour.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from ._lib import DnmfError
from .engine import Engine

DEFAULT_CUTOFF = 3.5          # per-axis box cutoff in units of sigma (1e-7 loss error)
DENSE_LIMIT_BYTES = 2 << 30   # largest dense A_t / pushforward array materialised on request


def _as_size(sz) -> list:
    return [int(v) for v in (sz.tolist() if torch.is_tensor(sz) else sz)]


class ExponentialFP(nn.Module):
    """Gaussian footprints resampled through a per-frame quadratic deformation
    (reference: Demix/dNMF.py:18-122)."""

    def __init__(self, sz, K, T, positions=None, shape_std=3, *, cutoff: float = DEFAULT_CUTOFF,
                 device=None, tiling: Optional[Sequence[int]] = None):
        super().__init__()
        size = _as_size(sz)
        self.engine = Engine(size, K, T, device)
        dev = self.engine.device
        self.K, self.T = int(K), int(T)
        self.cutoff = float(cutoff)
        # identity deformation per frame (Demix/dNMF.py:24-27); leaf tensor handed to the caller's optimiser
        beta = torch.zeros(10, 3, T, device=dev)
        beta[1, 0], beta[2, 1], beta[3, 2] = 1.0, 1.0, 1.0
        self.beta = beta.requires_grad_(True)
        self.sigma = (torch.ones(K) * shape_std).to(dev)                       # :29
        if positions is None:
            self.pos = (1 + torch.rand(K, 3) * torch.tensor(size)[None, :]).to(dev)   # :31
        else:
            self.pos = torch.as_tensor(positions).float().to(dev)              # :33
        self.sz = torch.as_tensor(size).to(dev)
        if tiling is not None:
            self.engine.set_tiling(*tiling)
        self.engine.set_footprints(self.pos, self.sigma, self.cutoff)
        self._A = None

    def set_footprints(self, positions=None, sigma=None, cutoff: Optional[float] = None):
        """Rebuild the per-axis tables after changing positions / widths / cutoff."""
        if positions is not None:
            self.pos = torch.as_tensor(positions).float().to(self.engine.device)
        if sigma is not None:
            self.sigma = torch.as_tensor(sigma).float().to(self.engine.device)
        if cutoff is not None:
            self.cutoff = float(cutoff)
        self.engine.set_footprints(self.pos, self.sigma, self.cutoff)
        self._A = None

    @property
    def flow_id(self) -> torch.Tensor:
        """Integer voxel coordinates [X,Y,Z,3] (Demix/dNMF.py:22), built on demand."""
        X, Y, Z = self.sz.tolist()
        dev = self.engine.device
        g = torch.meshgrid(torch.arange(X, device=dev), torch.arange(Y, device=dev), torch.arange(Z, device=dev),
                           indexing="ij")
        return torch.stack(g, 3).float()

    @property
    def transformed(self) -> torch.Tensor:
        return ExponentialFP.quadratic_basis(self.flow_id)

    @property
    def A(self) -> torch.Tensor:
        """Dense (untruncated) Gaussian volume [X,Y,Z,K] of Demix/dNMF.py:39-40.  The kernels never
        need it; it is materialised lazily for callers that read `fp.A` (demo.py:61)."""
        if self._A is None:
            X, Y, Z = self.sz.tolist()
            if X * Y * Z * self.K * 4 > DENSE_LIMIT_BYTES:
                raise DnmfError("fp.A would need %.1f GB; read the per-axis tables instead"
                                % (X * Y * Z * self.K * 4 / 2 ** 30))
            ax = [torch.arange(n, device=self.engine.device, dtype=torch.float32) for n in (X, Y, Z)]
            e = [-(ax[d][:, None] - self.pos[:, d][None, :]) ** 2 / self.sigma[None, :] ** 2 for d in range(3)]
            self._A = torch.exp(e[0][:, None, None, :] + e[1][None, :, None, :] + e[2][None, None, :, :])
        return self._A

    @staticmethod
    def quadratic_basis(P: torch.Tensor) -> torch.Tensor:
        """[..., 3] -> [..., 10] = [1,x,y,z,x2,y2,z2,xy,xz,yz] (Demix/dNMF.py:46-51)."""
        x, y, z = P[..., 0:1], P[..., 1:2], P[..., 2:3]
        return torch.cat((torch.ones_like(x), P, P * P, x * y, x * z, y * z), -1)

    def forward(self, times, C, want_dense: Optional[bool] = None):
        """Returns (A_tC[B,X,Y,Z], A_t[B,K,X,Y,Z], grid[X,Y,Z,3,B], reg[B]) like Demix/dNMF.py:53-62.
        The dense A_t / grid are produced only while they fit DENSE_LIMIT_BYTES (else None)."""
        times = [int(t) for t in (times.tolist() if hasattr(times, "tolist") else times)]
        B = len(times)
        X, Y, Z = self.sz.tolist()
        if want_dense is None:
            want_dense = B * self.K * X * Y * Z * 4 <= DENSE_LIMIT_BYTES
        C = torch.as_tensor(C).float().to(self.engine.device).contiguous()
        ids = torch.tensor(times, dtype=torch.int32)
        with torch.no_grad():
            AtC, At, grid = self.engine.forward(ids, self.beta.detach(), C, want_At=want_dense, want_grid=want_dense)
        reg = self.regularizer_values(times)
        return AtC, At, grid, reg

    def regularizer_values(self, times) -> torch.Tensor:
        """The detached diagnostic of Demix/dNMF.py:60-61 (log-det-Jacobian at two corners), on the CPU
        like the reference's torch.tensor([...]) (it carries no gradient)."""
        b = self.beta.detach()[:, :, list(times)].cpu()
        hi = (self.sz - 1).cpu().float()
        lo = torch.zeros(3)
        return torch.stack([ExponentialFP.log_det_jac(b[:, :, j], hi) ** 2 +
                            ExponentialFP.log_det_jac(b[:, :, j], lo) ** 2 for j in range(b.shape[2])])

    @staticmethod
    def log_det_jac(B, P):
        """log|det J| at point P with the reference's own cross-term indexing (Demix/dNMF.py:107-122)."""
        x, y, z = P[0], P[1], P[2]
        r = [(B[1, c] + 2 * B[4, c] * x + B[7, c] * y + B[9, c] * z,
              B[2, c] + 2 * B[5, c] * y + B[7, c] * x + B[8, c] * z,
              B[3, c] + 2 * B[6, c] * z + B[8, c] * y + B[9, c] * x) for c in range(3)]
        (a, b, c), (d, e, f), (g, h, i) = r
        return torch.log(abs(a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)))

    @staticmethod
    def log_det_jac_consistent(B, P):
        """log|det J_tau(P)| with cross-term indices that match quadratic_basis (xy -> row 7, xz -> row 8,
        yz -> row 9); differentiable in B.  Opt-in fix of the reference's inconsistency."""
        x, y, z = P[0], P[1], P[2]
        r = [(B[1, c] + 2 * B[4, c] * x + B[7, c] * y + B[8, c] * z,
              B[2, c] + 2 * B[5, c] * y + B[7, c] * x + B[9, c] * z,
              B[3, c] + 2 * B[6, c] * z + B[8, c] * x + B[9, c] * y) for c in range(3)]
        (a, b, c), (d, e, f), (g, h, i) = r
        return torch.log(abs(a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)))

    def jacobian_regularizer_grad(self, ids, gamma):
        """d/d beta of gamma * mean_t [log|det J_t(sz-1)|^2 + log|det J_t(0)|^2] for the batch frames:
        [10,3,B] on the device (30 numbers per frame; torch autograd, nothing voxel-sized)."""
        b = self.beta.detach()[:, :, ids].clone().requires_grad_(True)
        hi = (self.sz - 1).float()
        lo = torch.zeros(3, device=b.device)
        reg = ExponentialFP.log_det_jac_consistent(b, hi) ** 2 + ExponentialFP.log_det_jac_consistent(b, lo) ** 2
        (gamma * reg.mean()).backward()
        return b.grad

    @staticmethod
    def pushforward_chunks(dl, model, want_At: bool = True):
        """Lazy form of spatial_pushforward for volumes where the dense arrays cannot exist (825 GB at cfg2,
        Demix/dNMF.py:72): yields per batch of the loader (ids, A_t[B,K,X,Y,Z] or None, Y_i[B,X,Y,Z], Y[B,X,Y,Z]) as
        float32 CUDA tensors; nothing is kept between batches."""
        eng = model.fp.engine
        beta = model.fp.beta.detach()
        local = getattr(model, "_local_ids", lambda v: torch.as_tensor(v).to(torch.int32))
        for data in dl:
            ids = local(data[1])
            if data[0] is None:
                fd = eng.video()[ids.long().to(eng.device)].contiguous()
            else:
                fd = data[0].float().to(eng.device).contiguous()
            At = eng.forward(ids, beta, model.C, want_At=True)[1] if want_At else None
            yield ids, At, eng.iwarp(ids, beta, frames=fd), fd

    @staticmethod
    def max_projections(dl, model):
        """The projections demo.py:50-52 takes of the pushforward outputs, computed on the device without the dense
        arrays: (A_t.max(2) [X,Y,K,T'], Y_i.max(2) [X,Y,T'], Y.max(2) [X,Y,T']) as float32 CUDA tensors."""
        eng = model.fp.engine
        beta = model.fp.beta.detach()
        A_parts, Yi_parts, Y_parts = [], [], []
        for ids, _, yi, y in ExponentialFP.pushforward_chunks(dl, model, want_At=False):
            A_parts.append(eng.forward_maxz(ids, beta))        # [B,K,X,Y]
            Yi_parts.append(eng.frames_maxz(yi))               # [B,X,Y]
            Y_parts.append(eng.frames_maxz(y))
        A_max = torch.cat(A_parts, 0).permute(2, 3, 1, 0)
        return A_max, torch.cat(Yi_parts, 0).permute(1, 2, 0), torch.cat(Y_parts, 0).permute(1, 2, 0)

    @staticmethod
    def spatial_pushforward(dl, batch_size, sz, device, model):
        """Dense outputs of Demix/dNMF.py:69-93: (A_t[X,Y,Z,K,T'] f64, Y_i[X,Y,Z,T'] f64, Y[X,Y,Z,T'] f64)
        as numpy arrays, computed on the GPU batch by batch."""
        size = _as_size(sz)
        n = len(dl) * batch_size
        K = model.C.shape[0]
        need = size[0] * size[1] * size[2] * K * n * 8
        if need > 8 * DENSE_LIMIT_BYTES:
            raise DnmfError("spatial_pushforward would materialise %.1f GB; use update_footprints(dense=False)"
                            % (need / 2 ** 30))
        A_t = np.zeros((size[0], size[1], size[2], K, n))
        Y_i = np.zeros((size[0], size[1], size[2], n))
        Y = np.zeros((size[0], size[1], size[2], n))
        eng = model.fp.engine
        beta = model.fp.beta.detach()
        for bi, data in enumerate(dl):
            frames = data[0].float()
            ids = model._local_ids(data[1]) if hasattr(model, "_local_ids") else torch.as_tensor(data[1]).to(torch.int32)
            fd = frames.to(eng.device).contiguous()
            _, At, _ = eng.forward(ids, beta, model.C, want_At=True)
            yi = eng.iwarp(ids, beta, frames=fd)
            s = slice(bi * batch_size, bi * batch_size + fd.shape[0])
            A_t[..., s] = At.permute(2, 3, 4, 1, 0).double().cpu().numpy()
            Y[..., s] = fd.permute(1, 2, 3, 0).double().cpu().numpy()
            Y_i[..., s] = yi.permute(1, 2, 3, 0).double().cpu().numpy()
        return A_t, Y_i, Y


def loader_index_batches(loader):
    """The frame-id batches a `torch.utils.data.DataLoader` would yield, WITHOUT loading a frame.

    With an attached (HBM-resident) video the hot path needs only the ids of each minibatch, but walking the
    reference's DataLoader (`demo.py:34-35`) makes the dataset slice and collate every frame on the host: 2.75 GB of
    memcpy per epoch at cfg2 next to 3 ms of device work.  For datasets whose items are `(frame, own index)`
    (`Demix/dNMF.py:214-217`; marked `returns_frame_index`, or the reference's class names) the loader's batch sampler
    yields the same ids in the same order.  The random stream is consumed exactly as DataLoader iteration does
    (the iterator draws its base seed before the sampler seeds itself), so shuffled runs stay reproducible against
    the reference under the same `torch.manual_seed`.  Returns None when the shortcut does not apply."""
    from torch.utils.data import DataLoader
    if not isinstance(loader, DataLoader) or loader.batch_sampler is None:
        return None
    ds = loader.dataset
    if not (getattr(ds, "returns_frame_index", False) or
            type(ds).__name__ in ("SimulatedVideoDataset", "NeuroPALVideoDataset")):
        return None
    # map-style datasets with a finite batch sampler only (an IterableDataset's sampler never ends), and only for
    # the DataLoader internals this shortcut was checked against (tests/test_host.py compares it with real iteration)
    if getattr(loader, "_dataset_kind", 0) != 0 or not hasattr(loader.batch_sampler, "__len__"):
        return None
    if not _loader_shortcut_ok():
        return None
    offset = int(getattr(ds, "offset", 0))
    torch.empty((), dtype=torch.int64).random_(generator=loader.generator)      # _BaseDataLoaderIter._base_seed
    return [np.asarray(b, dtype=np.int64).reshape(-1) + offset for b in loader.batch_sampler]


_LOADER_SHORTCUT = None


def _loader_shortcut_ok() -> bool:
    """One self-check per process: on a tiny shuffled DataLoader, walking the batch sampler after one int64 draw
    must give the batches (and leave the global random stream in the state) that real iteration gives.  A torch
    release that changes how the iterator consumes the stream makes this False and the loaders are walked the
    ordinary way."""
    global _LOADER_SHORTCUT
    if _LOADER_SHORTCUT is None:
        from torch.utils.data import DataLoader, Dataset

        class _Probe(Dataset):
            def __len__(self):
                return 7

            def __getitem__(self, i):
                return torch.zeros(1), i

        state = torch.get_rng_state()
        try:
            torch.manual_seed(12345)
            real = [b[1].tolist() for b in DataLoader(_Probe(), batch_size=3, shuffle=True)]
            after_real = torch.rand(1).item()
            torch.manual_seed(12345)
            dl = DataLoader(_Probe(), batch_size=3, shuffle=True)
            torch.empty((), dtype=torch.int64).random_(generator=dl.generator)
            mine = [list(b) for b in dl.batch_sampler]
            after_mine = torch.rand(1).item()
            _LOADER_SHORTCUT = real == mine and after_real == after_mine
        except Exception:
            _LOADER_SHORTCUT = False
        finally:
            torch.set_rng_state(state)
    return _LOADER_SHORTCUT


def collect_id_batches(loader):
    """(ids int32[sum of batch sizes], offsets int32[nbatches + 1]) of one pass over `loader`: through the sampler
    when `loader_index_batches` applies, else from the second element of every item the loader yields."""
    batches = loader_index_batches(loader)
    if batches is None:
        batches = [np.asarray(data[1].cpu() if torch.is_tensor(data[1]) else data[1]).reshape(-1) for data in loader]
    offsets = np.zeros(len(batches) + 1, dtype=np.int32)
    if batches:
        offsets[1:] = np.cumsum([b.size for b in batches])
        ids = np.ascontiguousarray(np.concatenate(batches).astype(np.int32))
    else:
        ids = np.zeros(0, dtype=np.int32)
    return ids, offsets


class DeformableNMF:
    """dNMF fit: Adam on the per-frame deformation + multiplicative updates of the traces
    (reference: Demix/dNMF.py:124-194)."""

    def __init__(self, sz, K, T, positions=None, *, cutoff: float = DEFAULT_CUTOFF, deformation: str = "quadratic",
                 shape_std=3, device=None, tiling=None, verbose: bool = True, frame_offset: int = 0,
                 global_batch_scale: int = 1, jacobian_regularizer: bool = False):
        if deformation not in ("quadratic", "affine"):
            raise ValueError("deformation must be 'quadratic' or 'affine'")
        self.SpatialModel = ExponentialFP
        self.fp = ExponentialFP(sz=sz, K=K, T=T, positions=positions, shape_std=shape_std, cutoff=cutoff,
                                device=device, tiling=tiling)
        dev = self.fp.engine.device
        self.C = torch.rand((K, T)).to(dev)                                     # Demix/dNMF.py:130
        size = _as_size(sz)
        self.A = torch.rand((K, size[0], size[1])).to(dev)                      # :131 (unused, kept for parity)
        self._positions = None if positions is None else torch.as_tensor(positions).float()
        self._size = size
        self._D = None
        self.affine = deformation == "affine"
        self.fp.engine.set_affine(self.affine)
        # opt-in: make `gamma` of update_motion act (differentiable, index-consistent log-det-Jacobian
        # penalty).  OFF reproduces the reference, where the regulariser is a detached constant (F3).
        self.jacobian_regularizer = bool(jacobian_regularizer)
        self.verbose = verbose
        self.frame_offset = int(frame_offset)        # first global frame id of this rank's slab
        self.global_batch_scale = int(global_batch_scale)   # world size when every rank steps a local batch
        self.loss_history = []                       # python floats / CUDA scalars, one per Adam step
        self._loss_buf = None
        self._video_resident = False
        self._shared = None                          # extension state (enable_shared_learning)
        self.background = torch.zeros((), device=dev)

    # distance penalty of Demix/dNMF.py:133-135, only used by the (disabled) update_spatial: lazy
    @property
    def D(self):
        if self._positions is None:
            return None
        if self._D is None:
            X, Y, Z = self._size
            dev = self.fp.engine.device
            g = torch.stack(torch.meshgrid(torch.arange(X, device=dev), torch.arange(Y, device=dev),
                                           torch.arange(Z, device=dev), indexing="ij"), 3).reshape(-1, 3).double()
            d = torch.cdist(g, self._positions.to(dev).double())
            self._D = (1 - torch.exp(-.01 * d)).reshape(X, Y, Z, -1).cpu().numpy()
        return self._D

    # -- resident video (extension: keeps the frames in HBM so steps only send frame ids) ----------
    def attach_video(self, video: torch.Tensor, layout: str = "XYZT", copy: bool = True):
        """Upload this rank's frames once.  layout 'XYZT' is SimulatedVideoDataset.video
        (Demix/dNMF.py:203), 'TXYZ' is frame-major.  copy=False adopts a frame-major float32 CUDA tensor in place
        (no second slab in HBM; negatives are clamped in place like the reference's dataset does, :215)."""
        if layout == "XYZT":
            frames = video.permute(3, 0, 1, 2)
        elif layout == "TXYZ":
            frames = video
        else:
            raise ValueError("layout must be 'XYZT' or 'TXYZ'")
        if frames.shape[0] != self.fp.T:
            raise DnmfError("video has %d frames, model has T=%d" % (frames.shape[0], self.fp.T))
        if not copy:
            if not (frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous()):
                raise DnmfError("attach_video(copy=False) needs a contiguous frame-major float32 CUDA tensor")
            self.fp.engine.attach_frames(frames, clamp_negative=True)
            self._video_resident = True
            return
        self.fp.engine.upload_frames(frames.float().contiguous(), 0, clamp_negative=True)
        self._video_resident = True

    # -- EXTENSION (no reference counterpart, off by default) ---------------------------------------
    def enable_shared_learning(self, lr_pos=0.0, lr_sigma=0.0, lr_background=0.0, process_group=None,
                               device_step: bool = True, sigma_min: float = 0.5):
        """Also learn the SHARED parameters -- neuron positions, widths and a scalar background added to the
        model -- with Adam next to the per-frame deformation.  Their gradients come from
        dnmf_ext_loss_grad; with frames sharded over GPUs they are the only gradients that are all-reduced
        (one small NCCL all-reduce per iteration).  The reference keeps pos/sigma fixed and has no background
        (Demix/dNMF.py:29-33), so this has no reference oracle: it is tested against torch autograd.

        device_step=True (default): the whole iteration stays on the device (dnmf_ext_step_begin / _end): one packed
        gradient buffer for the all-reduce, Adam on pos / sigma / b with the library's kernel (betas and eps of the
        optimiser handed to update_motion), sigma clamped to >= sigma_min, ranges / tables / candidate lists rebuilt by
        kernels.  `fp.pos`, `fp.sigma`, `background` are refreshed from the device at the end of every update_motion
        call.  device_step=False keeps the first implementation (torch.optim.Adam on the three tensors and a table
        rebuild through the host per step); the two are tested against each other."""
        eng = self.fp.engine
        eng.ext_enable()
        self.fp.pos = self.fp.pos.detach().clone().requires_grad_(True)
        self.fp.sigma = self.fp.sigma.detach().clone().requires_grad_(True)
        self.background = self.background.detach().clone().requires_grad_(True)
        groups = [{"params": [self.fp.pos], "lr": lr_pos}, {"params": [self.fp.sigma], "lr": lr_sigma},
                  {"params": [self.background], "lr": lr_background}]
        self._shared = {"opt": torch.optim.Adam(groups), "group": process_group, "device_step": bool(device_step),
                        "lr": (float(lr_pos), float(lr_sigma), float(lr_background)), "sigma_min": float(sigma_min),
                        "packed": torch.zeros(4 * eng.K + 2, dtype=torch.float64, device=eng.device), "step": 0}
        eng.set_footprints(self.fp.pos, self.fp.sigma, self.fp.cutoff)
        if device_step:
            eng.ext_set_params(float(self.background.detach()), reset_adam_state=True)

    def _shared_step(self, ids, frames_dev, optimizer_group, st, step, B_global):
        """One iteration with shared-parameter learning: fused kernel (residual written) + parameter
        gradient kernel, all-reduce of the shared gradients, device Adam on beta, Adam on pos/sigma/b,
        table rebuild."""
        import torch.distributed as dist
        eng = self.fp.engine
        beta = self.fp.beta.detach()
        sh = self._shared
        if sh["device_step"]:
            packed, group = sh["packed"], sh["group"]
            eng.ext_step_begin(ids, beta, self.C, packed, frames=frames_dev, B_global=B_global)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(packed, group=group)         # the one collective of the extension
            if self._loss_buf is None:
                self._loss_buf = torch.zeros(1024, dtype=torch.float64, device=eng.device)
            slot = self._loss_buf[(step - 1) % 1024:(step - 1) % 1024 + 1]
            sh["step"] += 1
            eng.ext_step_end(beta, st["exp_avg"], st["exp_avg_sq"], optimizer_group["lr"], optimizer_group["betas"],
                             optimizer_group["eps"], step, self.affine, packed, B_global, *sh["lr"],
                             sigma_min=sh["sigma_min"], loss_out=slot)
            self.fp._A = None
            return slot
        grad = getattr(self, "_ext_grad", None)
        if grad is None:
            grad = self._ext_grad = torch.zeros_like(beta)
        _, sse, gpos, gsig, gbg = eng.ext_loss_grad(ids, beta, self.C, float(self.background.detach()), frames=frames_dev,
                                                    B_global=B_global, grad_beta=grad)
        loss = sse.sum() / (B_global * eng.N)
        group = self._shared["group"]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            flat = torch.cat((gpos.reshape(-1), gsig, gbg, loss.reshape(1)))
            dist.all_reduce(flat, group=group)       # the one collective of the extension
            K = eng.K
            gpos, gsig, gbg, loss = flat[:3 * K].reshape(K, 3), flat[3 * K:4 * K], flat[4 * K:4 * K + 1], flat[4 * K + 1]
        eng.adam_step(beta, grad, st["exp_avg"], st["exp_avg_sq"], optimizer_group["lr"], optimizer_group["betas"],
                      optimizer_group["eps"], step, self.affine)
        self.fp.pos.grad = gpos.float()
        self.fp.sigma.grad = gsig.float()
        self.background.grad = gbg.float().reshape(())
        self._shared["opt"].step()
        with torch.no_grad():
            self.fp.sigma.clamp_(min=0.5)
        eng.set_footprints(self.fp.pos, self.fp.sigma, self.fp.cutoff)
        self.fp._A = None
        return loss

    def _local_ids(self, ids) -> torch.Tensor:
        """Frame ids as the loader yields them (global ids when this rank's slab starts at `frame_offset`) ->
        int32 ids into the slab, validated on the host: an id outside the slab is an error, not a memory fault."""
        t = torch.as_tensor(ids).reshape(-1)
        if t.is_cuda:
            return (t - self.frame_offset).to(torch.int32) if self.frame_offset else t.to(torch.int32)
        t = t.to(torch.int64) - self.frame_offset
        if t.numel() and (int(t.min()) < 0 or int(t.max()) >= self.fp.T):
            bad = int(t.min()) if int(t.min()) < 0 else int(t.max())
            raise DnmfError("frame id %d is outside this model's slab [%d, %d)"
                            % (bad + self.frame_offset, self.frame_offset, self.frame_offset + self.fp.T))
        return t.to(torch.int32)

    # -- Adam state shared with the caller's optimiser ---------------------------------------------
    def _adam_state(self, optimizer):
        if not isinstance(optimizer, torch.optim.Adam) or type(optimizer) is not torch.optim.Adam:
            raise DnmfError("update_motion supports torch.optim.Adam only (demo.py:42), got %s" % type(optimizer).__name__)
        group = None
        for g in optimizer.param_groups:
            if any(p is self.fp.beta for p in g["params"]):
                group = g
        if group is None:
            raise DnmfError("the optimiser does not own dnmf.fp.beta")
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise DnmfError("weight_decay / amsgrad / maximize are not supported by the device Adam kernel")
        st = optimizer.state[self.fp.beta]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(self.fp.beta, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(self.fp.beta, memory_format=torch.preserve_format)
        return group, st

    def update_motion(self, dataloader, optimizer, gamma=0, epochs=20):
        """Demix/dNMF.py:181-194.  Each batch: fused forward/backward kernel + dense device Adam.
        `gamma` multiplies a detached constant in the reference and does not change
        the updates; it is accepted for signature compatibility."""
        eng = self.fp.engine
        group, st = self._adam_state(optimizer)
        beta = self.fp.beta.detach()
        # Resident video, reference-parity path: the frames the loader yields are not needed (attach_video), so the
        # whole epoch goes to the library in one call (dnmf_motion_epoch).  At the reference's batch size of 4
        # frames a step is ~12 us of device work; per-step Python and launch overhead would dominate.
        epoch_call = (self._video_resident and self._shared is None and not self.verbose and
                      not (self.jacobian_regularizer and gamma))   # verbose prints mid-epoch diagnostics per step
        for epoch in range(1, epochs + 1):
            if self.verbose:
                print("Epoch " + str(epoch))
            self.fp.train()
            if epoch_call:
                # one pass over the loader on the host: ids as numpy (no per-batch torch dispatch), one H2D copy;
                # a DataLoader over an index-returning dataset is walked through its sampler (no frame is loaded)
                ids_np, offsets = collect_id_batches(dataloader)
                nbatches = int(offsets.size) - 1
                if nbatches == 0:
                    continue
                ids_dev = self._local_ids(torch.from_numpy(ids_np)).to(eng.device)
                losses = torch.zeros(nbatches, dtype=torch.float64, device=eng.device)
                first = int(st["step"]) + 1
                eng.motion_epoch(ids_dev, offsets, beta, st["exp_avg"], st["exp_avg_sq"], self.C, group["lr"],
                                 group["betas"], group["eps"], first, self.affine,
                                 global_batch_scale=self.global_batch_scale, loss_out=losses)
                st["step"] += nbatches
                self.loss_history.extend(losses.unbind(0))
                continue
            for batch_idx, data in enumerate(dataloader):
                ids = self._local_ids(data[1])
                step = int(st["step"]) + 1
                lr, betas, eps = group["lr"], group["betas"], group["eps"]
                if self.jacobian_regularizer and gamma and self._shared is None:
                    fd = None if self._video_resident else data[0].float().to(eng.device).contiguous()
                    Bg = ids.numel() * self.global_batch_scale
                    idd = ids.to(eng.device)
                    grad, sse = eng.loss_grad(idd, beta, self.C, frames=fd, B_global=Bg)
                    grad[:, :, idd.long()] += self.fp.jacobian_regularizer_grad(idd.long(), float(gamma)) \
                        * (ids.numel() / Bg)
                    eng.adam_step(beta, grad, st["exp_avg"], st["exp_avg_sq"], lr, betas, eps, step, self.affine)
                    loss = sse.sum() / (Bg * eng.N)
                elif self._shared is not None:
                    fd = None if self._video_resident else data[0].float().to(eng.device).contiguous()
                    loss = self._shared_step(ids.to(eng.device), fd, group, st, step,
                                             ids.numel() * self.global_batch_scale)
                elif self._video_resident:
                    if self._loss_buf is None:
                        self._loss_buf = torch.zeros(1024, dtype=torch.float64, device=eng.device)
                    slot = self._loss_buf[(step - 1) % 1024:(step - 1) % 1024 + 1]
                    eng.motion_step(ids.to(eng.device), beta, st["exp_avg"], st["exp_avg_sq"], self.C, lr, betas, eps,
                                    step, self.affine, frames=None, B_global=ids.numel() * self.global_batch_scale,
                                    loss_out=slot)
                    loss = slot
                else:
                    frames = data[0]
                    if frames.is_cuda:
                        fd = frames.float().contiguous()
                        slot = torch.zeros(1, dtype=torch.float64, device=eng.device)
                        eng.motion_step(ids.to(eng.device), beta, st["exp_avg"], st["exp_avg_sq"], self.C, lr, betas,
                                        eps, step, self.affine, frames=fd,
                                        B_global=ids.numel() * self.global_batch_scale, loss_out=slot)
                        loss = slot
                    else:
                        fh = frames.float().contiguous()   # clamping at 0 is the dataset's job (Demix/dNMF.py:215)
                        loss = eng.motion_step_host(fh, ids.contiguous(), beta, st["exp_avg"], st["exp_avg_sq"],
                                                    self.C, lr, betas, eps, step, self.affine,
                                                    B_global=ids.numel() * self.global_batch_scale)
                st["step"] += 1
                self.loss_history.append(loss if isinstance(loss, float) else loss.clone())
                if self.verbose and batch_idx % 10 == 0:
                    print("Recon: " + str(float(loss)))
                    print("Reg: " + str(self.fp.regularizer_values(ids.tolist())))
        if self._shared is not None and self._shared["device_step"]:
            self._refresh_shared_params()
        eng.check_status()      # surfaces an error an asynchronous step found on the device (one sync per call)

    def _refresh_shared_params(self):
        """fp.pos / fp.sigma / background <- what the device-resident shared step has made of them."""
        pos, sigma, bg = self.fp.engine.ext_get_params()
        with torch.no_grad():
            self.fp.pos.copy_(pos)
            self.fp.sigma.copy_(sigma)
            self.background.copy_(bg)

    def losses(self) -> np.ndarray:
        """Per-step reconstruction losses recorded by update_motion (one device sync)."""
        hist = self.loss_history
        if hist and all(torch.is_tensor(l) for l in hist):
            return torch.stack([l.reshape(()) for l in hist]).double().cpu().numpy()   # one D2H copy
        return np.asarray([float(l) for l in hist])

    # -- traces -------------------------------------------------------------------------------------
    @staticmethod
    def update_temporal(A_t, C, Y, gamma=None, device=None):
        """Stand-alone multiplicative update on dense arrays (Demix/dNMF.py:139-149): fp64 kernels on the GPU
        (`dnmf_update_temporal_dense`), numpy array back like the reference."""
        from .engine import update_temporal_dense
        return update_temporal_dense(A_t, C, Y, gamma=gamma, device=device).cpu().numpy()

    @staticmethod
    def update_spatial(A, C, Y_i, D=None, gamma=None, device=None):
        """Non-parametric footprint update of Demix/dNMF.py:151-160 (never called by the reference's fit loop,
        :174): one fused fp64 pass over the registered video (`dnmf_update_spatial`), numpy array back."""
        from .engine import update_spatial_dense
        return update_spatial_dense(A, C, Y_i, D=D, gamma=gamma, device=device).cpu().numpy()

    def update_spatial_on_grid(self, A, Y_i, gamma=1e0):
        """update_spatial for this model's traces with the distance penalty of Demix/dNMF.py:133-135 computed on the
        fly from the neuron positions (the reference stores it as a dense fp64 [X,Y,Z,K] array, 826 MB at cfg2):
        A[X,Y,Z,K], Y_i[X,Y,Z,T] -> the updated A as a CUDA fp64 tensor."""
        from .engine import update_spatial_dense
        if self._positions is None:
            raise DnmfError("update_spatial_on_grid needs the positions the model was built with")
        return update_spatial_dense(A, self.C, Y_i, gamma=gamma, positions=self._positions, grid=self._size,
                                    device=self.fp.engine.device)

    def update_traces(self, testloader=None, gamma_c=1e-2, iter_c=10, halo_exchange=None):
        """Device-only trace update: statistics kernel over all frames, then iter_c sweeps.
        `halo_exchange(first, last) -> (prev, next)` lets a multi-GPU caller swap boundary columns."""
        eng = self.fp.engine
        beta = self.fp.beta.detach()
        if self._video_resident:
            # attached video: the loader is walked for its frame ids only (no host frames cross the link again)
            if testloader is None:
                ids = torch.arange(self.fp.T, dtype=torch.int32)
            else:
                ids = self._local_ids(torch.from_numpy(collect_id_batches(testloader)[0]))
            ids = ids.to(eng.device)
            step = 512
            for i in range(0, int(ids.numel()), step):
                eng.mu_stats(ids[i:i + step], beta)
        else:
            for data in testloader:
                ids = self._local_ids(data[1])
                fd = data[0].float().to(eng.device).contiguous()
                eng.mu_stats(ids, beta, frames=fd)
        if halo_exchange is None:
            eng.mu_sweeps(self.C, gamma_c, iter_c)
        else:
            eng.mu_begin(self.C)
            for _ in range(iter_c):
                first, last = eng.mu_boundary()
                prev, nxt = halo_exchange(first, last)
                eng.mu_sweep(gamma_c, prev, nxt)
            eng.mu_end(self.C)

    def update_footprints(self, testloader, batch_size, sz, gamma_c=1e-2, gamma_a=1e0, iter_c=10, dense=None):
        """Demix/dNMF.py:163-179: pushforward, then iter_c multiplicative updates of C.
        Returns (A_t, Y_i, Y) as numpy fp64 arrays like the reference while they fit in memory
        (dense=None decides by size; dense=False skips them and returns (None, None, None))."""
        size = _as_size(sz)
        n = len(testloader) * batch_size
        if dense is None:
            dense = size[0] * size[1] * size[2] * self.C.shape[0] * n * 8 <= 8 * DENSE_LIMIT_BYTES
        out = (None, None, None)
        if dense:
            with torch.no_grad():
                out = self.SpatialModel.spatial_pushforward(testloader, batch_size, sz, self.fp.engine.device, self)
        self.update_traces(testloader, gamma_c=gamma_c, iter_c=iter_c)
        return out

    def fit(self, dataloader, testloader, optimizer, batch_size, sz, outer=5, epochs=10, iter_c=50, gamma=1,
            gamma_c=0, dense=False):
        """The schedule of demo.py:44-46."""
        out = None
        for _ in range(outer):
            self.update_motion(dataloader, optimizer, gamma=gamma, epochs=epochs)
            out = self.update_footprints(testloader, batch_size, sz, gamma_c=gamma_c, iter_c=iter_c, dense=dense)
        return out
