"""EXTENSION tests (not reference parity): shared-parameter gradients (positions, widths, scalar
background) against torch autograd over the extended restatement oracle.ExtendedPort."""
import numpy as np
import pytest
import torch

from oracle import dnmf_oracle as O

pytestmark = pytest.mark.gpu


def _setup(sz, K, T, seed, tiling=None):
    from dnmf_b200.engine import Engine
    rng = np.random.default_rng(seed)
    pos = (rng.random((K, 3)) * (np.asarray(sz) - 1)).astype(np.float32)
    sig = (2.0 + rng.random(K)).astype(np.float32)
    g = torch.Generator().manual_seed(seed)
    s = torch.tensor([.5, .02, .02, .02, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4])[:, None, None]
    beta = O.identity_beta(T) + s * torch.randn(10, 3, T, generator=g)
    C = torch.rand(K, T, generator=g)
    frames = torch.rand(T, *sz, generator=g)
    e = Engine(sz, K, T)
    if tiling:
        e.set_tiling(*tiling)
    e.ext_enable()
    e.set_footprints(pos, sig, 0.0)
    return e, pos, sig, beta, C, frames


@pytest.mark.parametrize("tiling", [None, (1, 1, 0, 0, 1), (2, 2, 0, 0, 1), (1, 1, 0, 0, 2, 2)])
def test_shared_parameter_gradients_vs_autograd(tiling):
    sz, K, T = [20, 14, 5], 6, 4
    e, pos, sig, beta, C, frames = _setup(sz, K, T, 3, tiling)
    bg = 0.07
    port = O.ExtendedPort(sz, pos, sig, C, beta, bg)
    loss, gbeta, gpos, gsig, gbg = port.grads(frames, list(range(T)))
    gb, sse, dpos, dsig, dbg = e.ext_loss_grad(torch.arange(T), beta.cuda(), C.cuda(), bg, frames=frames.cuda())
    N = int(np.prod(sz))
    assert abs(float(sse.sum()) / (T * N) - loss) <= 1e-5 * loss
    assert np.abs(gb.cpu().numpy() - gbeta).max() <= 3e-5 * np.abs(gbeta).max()
    assert np.abs(dpos.cpu().numpy() - gpos).max() <= 2e-4 * np.abs(gpos).max()
    assert np.abs(dsig.cpu().numpy() - gsig).max() <= 2e-4 * np.abs(gsig).max()
    assert abs(float(dbg) - gbg) <= 1e-4 * abs(gbg)


def test_shared_gradients_sum_over_frame_shards():
    """frame sharding: the shared-parameter gradients of two half-batches (B_global = all frames) add up to
    the full-batch gradients -- the quantity a multi-GPU run all-reduces."""
    sz, K, T = [20, 14, 5], 6, 6
    e, pos, sig, beta, C, frames = _setup(sz, K, T, 5)
    b, c, f = beta.cuda(), C.cuda(), frames.cuda()
    _, sse, dpos, dsig, dbg = e.ext_loss_grad(torch.arange(T), b, c, 0.0, frames=f)
    ids0, ids1 = torch.arange(0, 3), torch.arange(3, 6)
    _, s0, p0, g0, b0 = e.ext_loss_grad(ids0, b, c, 0.0, frames=f[:3].contiguous(), B_global=T)
    _, s1, p1, g1, b1 = e.ext_loss_grad(ids1, b, c, 0.0, frames=f[3:].contiguous(), B_global=T)
    np.testing.assert_allclose((p0 + p1).cpu().numpy(), dpos.cpu().numpy(), rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose((g0 + g1).cpu().numpy(), dsig.cpu().numpy(), rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(float(b0 + b1), float(dbg), rtol=1e-9)
    np.testing.assert_allclose(torch.cat((s0, s1)).cpu().numpy(), sse.cpu().numpy(), rtol=0, atol=0)


def test_learning_positions_reduces_loss():
    """a few Adam steps on beta, positions, widths and background through the public API (flag ON)."""
    from dnmf_b200 import DeformableNMF, FrameDataset
    from torch.utils.data import DataLoader
    torch.manual_seed(0)
    sz, K, T = [24, 16, 4], 5, 8
    rng = np.random.default_rng(1)
    true_pos = (rng.random((K, 3)) * (np.asarray(sz) - 1)).astype(np.float32)
    gen = DeformableNMF(sz, K, T, positions=torch.tensor(true_pos), cutoff=0.0, verbose=False)
    frames = gen.fp(list(range(T)), gen.C)[0].cpu() + 0.05            # model frames + background 0.05
    start = torch.tensor(true_pos + rng.normal(0, 0.7, true_pos.shape).astype(np.float32))
    dn = DeformableNMF(sz, K, T, positions=start, cutoff=0.0, verbose=False)
    dn.C = gen.C.clone()
    dn.enable_shared_learning(lr_pos=0.05, lr_sigma=0.01, lr_background=0.01)
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    loader = DataLoader(FrameDataset(frames), batch_size=T, shuffle=False)
    dn.update_motion(loader, opt, epochs=40)
    losses = dn.losses()
    assert losses[-1] < 0.25 * losses[0]
    err0 = np.abs(start.numpy() - true_pos).mean()
    err1 = np.abs(dn.fp.pos.detach().cpu().numpy() - true_pos).mean()
    assert err1 < 0.6 * err0
    assert abs(float(dn.background.detach()) - 0.05) < 0.04


def test_device_resident_shared_step_matches_the_torch_optimiser_route():
    """(extension, not reference parity)  dnmf_ext_step_begin / _end -- device Adam on pos / sigma / b, tables and
    candidate lists rebuilt by kernels, no host round trip -- against the first implementation (torch.optim.Adam on
    the three tensors, dnmf_set_footprints through the host every step): same parameters after 6 iterations."""
    from dnmf_b200 import DeformableNMF, FrameDataset
    from torch.utils.data import DataLoader
    sz, K, T = [28, 20, 5], 6, 8
    rng = np.random.default_rng(3)
    pos = (rng.random((K, 3)) * (np.asarray(sz) - 1)).astype(np.float32)
    frames = torch.tensor(rng.random((T, *sz)).astype(np.float32))
    C0 = torch.tensor(rng.random((K, T)).astype(np.float32))
    out = []
    for device_step in (True, False):
        dn = DeformableNMF(sz, K, T, positions=torch.tensor(pos), cutoff=3.5, verbose=False)
        dn.C = C0.clone().cuda()
        dn.enable_shared_learning(lr_pos=0.03, lr_sigma=0.01, lr_background=0.01, device_step=device_step)
        opt = torch.optim.Adam([dn.fp.beta], lr=1e-4)
        loader = DataLoader(FrameDataset(frames), batch_size=4, shuffle=False)
        dn.update_motion(loader, opt, epochs=3)
        out.append((dn.fp.pos.detach().cpu().numpy(), dn.fp.sigma.detach().cpu().numpy(), float(dn.background.detach()),
                    dn.fp.beta.detach().cpu().numpy(), dn.losses()))
        # the context's tables follow the learned parameters: a fresh context built from them gives the same forward
        ref = DeformableNMF(sz, K, T, positions=dn.fp.pos.detach().cpu(), cutoff=3.5, verbose=False,
                            shape_std=dn.fp.sigma.detach().cpu())
        with torch.no_grad():
            ref.fp.beta.copy_(dn.fp.beta)
        y0 = dn.fp([0, 3], dn.C)[0]
        y1 = ref.fp([0, 3], dn.C)[0]
        assert float((y0 - y1).abs().max()) <= 1e-6 * float(y1.abs().max())
    a, b = out
    assert np.abs(a[0] - pos).max() > 1e-3                                 # the positions did move
    np.testing.assert_allclose(a[0], b[0], rtol=0, atol=2e-5)
    np.testing.assert_allclose(a[1], b[1], rtol=0, atol=2e-5)
    assert abs(a[2] - b[2]) < 2e-5
    np.testing.assert_allclose(a[3], b[3], rtol=0, atol=2e-6)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-5)


def test_two_rank_shared_learning_matches_single_rank(tmp_path):
    """2 processes (one GPU, gloo), each owning half the frames, all-reduce the shared gradients every
    iteration: positions / widths / background end up equal to the single-process full-batch run."""
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worker = tmp_path / "w.py"
    worker.write_text(r'''
import sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dnmf_b200 import DeformableNMF, FrameDataset
from dnmf_b200.sharding import frame_slab
from torch.utils.data import DataLoader
world = int(sys.argv[4])
if world > 1:
    dist.init_process_group("gloo")
rank = dist.get_rank() if world > 1 else 0
torch.cuda.set_device(0)
d = np.load(sys.argv[2])
sz, K, T = d["sz"].tolist(), int(d["K"]), int(d["T"])
start, count = frame_slab(T, world, rank)
dn = DeformableNMF(sz, K, count, positions=torch.tensor(d["pos"]), cutoff=0.0, verbose=False, global_batch_scale=world)
dn.C = torch.tensor(d["C"][:, start:start + count]).contiguous().cuda()
dn.enable_shared_learning(lr_pos=0.02, lr_sigma=0.01, lr_background=0.01)
opt = torch.optim.Adam([dn.fp.beta], lr=1e-4)
loader = DataLoader(FrameDataset(torch.tensor(d["frames"][start:start + count])), batch_size=count, shuffle=False)
dn.update_motion(loader, opt, epochs=5)
np.savez(sys.argv[3] + "_%d_%d.npz" % (world, rank), pos=dn.fp.pos.detach().cpu().numpy(),
         sigma=dn.fp.sigma.detach().cpu().numpy(), bg=float(dn.background.detach()), losses=dn.losses(),
         beta=dn.fp.beta.detach().cpu().numpy())
if world > 1:
    dist.destroy_process_group()
''')
    sz, K, T = [20, 14, 4], 4, 6
    rng = np.random.default_rng(8)
    pos = (rng.random((K, 3)) * (np.asarray(sz) - 1)).astype(np.float32)
    data = tmp_path / "d.npz"
    np.savez(data, sz=np.asarray(sz), K=K, T=T, pos=pos, C=rng.random((K, T)).astype(np.float32),
             frames=rng.random((T, *sz)).astype(np.float32))
    out = str(tmp_path / "o")
    res = subprocess.run([sys.executable, str(worker), root, str(data), out, "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(worker), root, str(data), out,
                          "2"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    one = np.load(out + "_1_0.npz")
    two = [np.load(out + "_2_%d.npz" % r) for r in range(2)]
    for t in two:
        np.testing.assert_allclose(t["pos"], one["pos"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(t["sigma"], one["sigma"], rtol=0, atol=2e-5)
        assert abs(float(t["bg"]) - float(one["bg"])) < 2e-5
        np.testing.assert_allclose(t["losses"], one["losses"], rtol=1e-5)
    np.testing.assert_allclose(np.concatenate([t["beta"] for t in two], 2), one["beta"], rtol=0, atol=2e-6)
    assert np.abs(one["pos"] - pos).max() > 1e-3          # the shared parameters did move
