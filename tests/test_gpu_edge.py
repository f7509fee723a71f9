"""GPU edge cases and size-independent properties of the CUDA path (through the C ABI)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import dnmf_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(sz, K, T, seed, sigma=2.5, beta_scale=1.0):
    rng = np.random.default_rng(seed)
    pos = (rng.random((K, 3)) * np.asarray(sz)).astype(np.float32)
    sig = np.full(K, sigma, np.float32)
    g = torch.Generator().manual_seed(seed)
    s = torch.tensor([1.0, .02, .02, .02, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4])[:, None, None] * beta_scale
    beta = O.identity_beta(T) + s * torch.randn(10, 3, T, generator=g)
    beta[:, :, 0] = O.identity_beta(1)[:, :, 0]
    C = torch.rand(K, T, generator=g)
    frames = torch.rand(T, *sz, generator=g)
    return pos, sig, beta, C, frames


def _check(sz, K, T, seed, cutoff, tiling, sigma=2.5, beta_scale=1.0, tol=3e-5):
    from dnmf_b200.engine import Engine
    pos, sig, beta, C, frames = _case(sz, K, T, seed, sigma, beta_scale)
    e = Engine(sz, K, T)
    e.set_tiling(*tiling)
    e.set_footprints(pos, sig, cutoff)
    grad, sse = e.loss_grad(torch.arange(T), beta.cuda(), C.cuda(), frames=frames.cuda())
    tabs, _ = O.axis_tables(pos, sig, sz, cutoff)
    loss, gref = O.closed_form_step(frames.numpy(), list(range(T)), beta.numpy(), C.numpy(), tabs, sz)
    N = int(np.prod(sz))
    assert abs(float(sse.sum()) / (T * N) - loss) <= 1e-5 * loss
    err = np.abs(grad.cpu().numpy() - gref).max() / np.abs(gref).max()
    assert err < tol, err
    return e


@pytest.mark.parametrize("sz", [[13, 7, 5], [8, 4, 1], [33, 9, 2], [17, 30, 11], [50, 50, 2]])
@pytest.mark.parametrize("tiling", [(1, 1, 0, 0), (2, 2, 0, 0), (1, 1, 0, 0, 2), (2, 1, 0, 0, 2), (2, 2, 0, 0, 2),
                                    (1, 1, 0, 0, 2, 2), (1, 1, 0, 0, 2, 4)])
def test_ragged_sizes(sz, tiling):
    """volumes that do not divide into tiles, odd depths (no bulk-copy alignment), singleton z; z-split warp groups
    (warps_z = 2, 4) with fewer z planes than warps."""
    _check(sz, 4, 3, seed=sum(sz), cutoff=3.5, tiling=tiling)


def test_depth_chunked_tiles():
    _check([20, 12, 40], 6, 2, seed=5, cutoff=3.0, tiling=(1, 1, 16, 0))
    _check([20, 12, 40], 6, 2, seed=5, cutoff=3.0, tiling=(2, 1, 7, 0))


def test_list_longer_than_staged_capacity():
    """more listed neurons than staged slots: the overflow path through the global tables."""
    e = _check([24, 16, 6], 40, 2, seed=9, cutoff=0.0, tiling=(1, 1, 0, 4), sigma=4.0)
    assert e.tiling()["cap"] <= 6
    _check([24, 16, 6], 40, 2, seed=9, cutoff=0.0, tiling=(2, 4, 0, 8), sigma=4.0)


def test_window_wider_than_staged_slices():
    """a strong deformation makes the tile window outgrow the staged slices: global-table path."""
    _check([24, 16, 6], 8, 3, seed=11, cutoff=3.5, tiling=(1, 1, 0, 0), beta_scale=8.0, tol=1e-4)


def test_neurons_outside_volume_and_empty_ranges():
    from dnmf_b200.engine import Engine
    sz = [16, 12, 4]
    pos = np.array([[-40., 5., 1.], [8., 6., 2.], [100., 100., 100.]], np.float32)
    sig = np.full(3, 2.0, np.float32)
    e = Engine(sz, 3, 2)
    e.set_footprints(pos, sig, 3.0)
    rng = e.ranges()
    assert np.array_equal(rng, O.axis_ranges(pos, sig, sz, 3.0))
    assert rng[0, 0, 0] > rng[0, 0, 1] and rng[2, 1, 0] > rng[2, 1, 1]       # empty ranges
    beta = O.identity_beta(2).cuda()
    counts, offsets, ids, _ = e.bin_tiles(beta, torch.arange(2))
    assert set(ids.tolist()) <= {1}
    C = torch.ones(3, 2).cuda()
    frames = torch.rand(2, *sz)
    grad, sse = e.loss_grad(torch.arange(2), beta, C, frames=frames.cuda())
    tabs, _ = O.axis_tables(pos, sig, sz, 3.0)
    loss, gref = O.closed_form_step(frames.numpy(), [0, 1], beta.cpu().numpy(), C.cpu().numpy(), tabs, sz)
    assert abs(float(sse.sum()) / (2 * 768) - loss) <= 1e-6 * loss


@pytest.mark.parametrize("sz,tiling", [([40, 24, 9], (1, 1, 0, 0, 2)), ([33, 18, 32], (1, 1, 0, 0, 2)),
                                       ([40, 24, 9], (2, 2, 0, 0, 2)), ([24, 20, 5], (2, 2, 0, 0, 1))])
def test_affine_main_loops_bit_equal(sz, tiling):
    """dnmf_set_affine: affine frames take main loops without the z^2 Horner term and the z^2 gradient moments.
    Same arithmetic per voxel; the affine instantiation reduces 16 instead of 32 values per tile-frame, so the lane
    sums are taken in another order: rows 0..3 and the SSE agree to fp32 summation noise, rows 4..9 are zero.  A frame
    with quadratic coefficients takes the generic loop; its rows 4..9 are still returned as zero."""
    from dnmf_b200.engine import Engine
    K, T = 12, 6
    pos, sig, beta, C, frames = _case(sz, K, T, seed=21)
    beta[4:, :, :T - 1] = 0.0                      # frames 0..T-2 affine, the last one keeps its quadratic terms
    e = Engine(sz, K, T)
    e.set_tiling(*tiling)
    e.set_footprints(pos, sig, 3.5)
    ids = torch.arange(T)
    g0, s0 = e.loss_grad(ids, beta.cuda(), C.cuda(), frames=frames.cuda())
    e.set_affine(True)
    g1, s1 = e.loss_grad(ids, beta.cuda(), C.cuda(), frames=frames.cuda())
    e.set_affine(False)
    g2, s2 = e.loss_grad(ids, beta.cuda(), C.cuda(), frames=frames.cuda())
    assert torch.equal(s0, s2) and torch.equal(g0, g2)
    assert float(((s0 - s1).abs() / s0).max()) < 1e-6
    assert float((g0[:4] - g1[:4]).abs().max() / g0[:4].abs().max()) < 2e-6
    assert float(g1[4:].abs().max()) == 0.0 and float(g0[4:].abs().max()) > 0.0
    # the step calls do the same for the duration of a call made with affine != 0
    m, v = torch.zeros(10, 3, T).cuda(), torch.zeros(10, 3, T).cuda()
    b_a, b_b = beta.clone().cuda(), beta.clone().cuda()
    e.motion_step(ids.int().cuda(), b_a, m, v, C.cuda(), 1e-3, (0.9, 0.999), 1e-8, 1, True, frames=frames.cuda())
    gz = g1.clone()
    e.adam_step(b_b, gz, torch.zeros_like(m), torch.zeros_like(v), 1e-3, (0.9, 0.999), 1e-8, 1, affine=True)
    assert torch.equal(b_a, b_b)


def test_affine_freezes_quadratic_rows():
    from dnmf_b200.engine import Engine
    e = Engine([8, 8, 2], 2, 3)
    p = O.identity_beta(3).cuda()
    g = torch.ones(10, 3, 3).cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    e.adam_step(p, g, m, v, 1e-2, (0.9, 0.999), 1e-8, 1, affine=True)
    d = (p - O.identity_beta(3).cuda()).abs()
    assert float(d[4:].max()) == 0.0 and float(d[:4].min()) > 0.0
    assert float(m[4:].abs().max()) == 0.0 and float(v[4:].abs().max()) == 0.0


def test_bitwise_reproducible_and_frame_order_independent():
    """fixed-order reductions: same inputs -> identical bits; a frame's gradient does not depend on
    which other frames share its batch or on its slot."""
    from dnmf_b200.engine import Engine
    sz, K, T = [40, 24, 9], 12, 6
    pos, sig, beta, C, frames = _case(sz, K, T, 21)
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    e.upload_frames(frames, clamp_negative=False)
    b, c = beta.cuda(), C.cuda()
    g1, s1 = e.loss_grad(torch.arange(T), b, c)
    g2, s2 = e.loss_grad(torch.arange(T), b, c)
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    perm = torch.tensor([4, 1, 5])
    g3, s3 = e.loss_grad(perm, b, c, B_global=T)
    assert torch.equal(g3[:, :, perm], g1[:, :, perm]) and torch.equal(s3, s1[perm])
    assert float(g3[:, :, [0, 2, 3]].abs().max()) == 0.0


def test_forward_is_linear_in_traces_and_zero_for_zero_traces():
    from dnmf_b200.engine import Engine
    sz, K, T = [32, 20, 7], 9, 2
    pos, sig, beta, C, _ = _case(sz, K, T, 33)
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    b = beta.cuda()
    y1, _, _ = e.forward(torch.arange(T), b, C.cuda())
    y2, _, _ = e.forward(torch.arange(T), b, (2 * C).cuda())
    y0, _, _ = e.forward(torch.arange(T), b, torch.zeros_like(C).cuda())
    assert float(y0.abs().max()) == 0.0
    np.testing.assert_allclose(y2.cpu().numpy(), 2 * y1.cpu().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tiling", [None, (1, 1, 0, 0, 2, 2), (1, 1, 0, 0, 2, 4), (2, 2, 0, 0, 2)])
def test_perfect_model_has_zero_loss_and_gradient(tiling):
    """Y := model output  =>  sse == 0 and gradient == 0 exactly (encode -> decode round trip): the forward-only and the
    fit instantiations round Yhat identically, also in the z-split layouts (shared x / z slice loads)."""
    from dnmf_b200.engine import Engine
    sz, K, T = [32, 20, 7], 9, 3
    pos, sig, beta, C, _ = _case(sz, K, T, 35)
    e = Engine(sz, K, T)
    if tiling:
        e.set_tiling(*tiling)
    e.set_footprints(pos, sig, 3.5)
    b, c = beta.cuda(), C.cuda()
    y, _, _ = e.forward(torch.arange(T), b, c)
    grad, sse = e.loss_grad(torch.arange(T), b, c, frames=y.contiguous())
    assert float(sse.abs().max()) == 0.0 and float(grad.abs().max()) == 0.0


def test_iwarp_matches_scipy_nearest():
    """registered video Y_i (Demix/dNMF.py:81-83,95-103) against scipy's NearestNDInterpolator on a
    deformation without exact ties."""
    import scipy.interpolate
    from dnmf_b200.engine import Engine
    sz, K, T = [18, 14, 5], 3, 2
    pos, sig, beta, C, frames = _case(sz, K, T, 41, beta_scale=0.7)
    beta[:, :, 0] = beta[:, :, 1] * 1.01
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    out = e.iwarp(torch.arange(T), beta.cuda(), frames=frames.cuda()).cpu().numpy()
    grid = np.array(np.where(np.ones(sz))).T
    for t in range(T):
        u = O.sample_coords  # noqa: F841  (coordinates below follow the reference: scale by sz, not sz-1)
        _, phi = O.voxel_basis(sz)
        q = torch.einsum("mnza,ab->mnzb", phi, beta[:, :, t])
        un = 2 * q / (torch.tensor(sz) - 1) - 1
        f = ((un + 1) / 2) * torch.tensor(sz).float()
        interp = scipy.interpolate.NearestNDInterpolator(f.reshape(-1, 3).numpy(), frames[t].reshape(-1).numpy())
        ref = interp(grid).reshape(sz)
        agree = (out[t] == ref).mean()
        assert agree > 0.995, agree


def test_full_size_cfg2_properties():
    """BASELINE.json config 2 shape (256x128x21, K=150): determinism, perfect-model round trip and the
    zero-trace checksum at full size (the oracle is too slow here, properties are size-independent)."""
    from dnmf_b200.engine import Engine
    sz, K, T = [256, 128, 21], 150, 4
    pos, sig, beta, C, _ = _case(sz, K, T, 51, sigma=3.0, beta_scale=0.2)
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    b, c = beta.cuda(), C.cuda()
    y, _, _ = e.forward(torch.arange(T), b, c)
    assert torch.isfinite(y).all() and float(y.max()) > 0
    g0, s0 = e.loss_grad(torch.arange(T), b, c, frames=y.contiguous())
    assert float(s0.abs().max()) == 0.0 and float(g0.abs().max()) == 0.0
    frames = (y * 1.1).contiguous()
    g1, s1 = e.loss_grad(torch.arange(T), b, c, frames=frames)
    g2, s2 = e.loss_grad(torch.arange(T), b, c, frames=frames)
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    # sse of Y = 1.1*Yhat is 0.01 * sum(Yhat^2): checksum of checksums
    ref = 0.01 * (y.double() ** 2).sum(dim=(1, 2, 3))
    np.testing.assert_allclose(s1.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4)
    # a slice of the full-size volume against the oracle's closed form (one frame, one x-slab)
    tabs, _ = O.axis_tables(pos, sig, sz, 3.5)
    yh, _, _ = O.closed_form_frame(np.zeros(sz, np.float32), beta[:, :, 1].numpy(), C[:, 1].numpy(), tabs, sz)
    # q is a 10-term fp32 sum of magnitude ~255: summation order (Horner+FMA here, sequential in the
    # oracle, MKL bmm in the reference) moves samples by a few ulp(255) = 3e-5 px -> up to ~1e-5 in A_tC
    np.testing.assert_allclose(y[1].cpu().numpy(), yh, atol=3e-5)


def test_two_rank_frame_sharding_matches_single_rank(tmp_path):
    """2 processes on ONE GPU, each owning half the frames (B_global = all frames): the concatenated
    beta and the all-reduced loss equal the single-process result bit for bit."""
    worker = tmp_path / "w.py"
    worker.write_text(r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dnmf_b200.engine import Engine
from dnmf_b200.sharding import frame_slab, allreduce_loss
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(0)
d = np.load(sys.argv[2])
sz, K, T = d["sz"].tolist(), int(d["K"]), int(d["T"])
start, count = frame_slab(T, world, rank)
e = Engine(sz, K, count)
e.set_footprints(d["pos"], d["sig"], 3.5)
e.upload_frames(torch.tensor(d["frames"][start:start + count]), clamp_negative=False)
beta = torch.tensor(d["beta"][:, :, start:start + count]).contiguous().cuda()
C = torch.tensor(d["C"][:, start:start + count]).contiguous().cuda()
m, v = torch.zeros_like(beta), torch.zeros_like(beta)
loss = torch.zeros(1, dtype=torch.float64, device="cuda")
ids = torch.arange(count, dtype=torch.int32, device="cuda")
losses = []
for step in range(1, 4):
    e.motion_step(ids, beta, m, v, C, 1e-3, (0.9, 0.999), 1e-8, step, False, B_global=T, loss_out=loss)
    l = loss.cpu().clone()
    allreduce_loss(l)
    losses.append(float(l))
np.savez(sys.argv[3] + "_%d.npz" % rank, beta=beta.cpu().numpy(), losses=np.asarray(losses), start=start)
dist.destroy_process_group()
''')
    from dnmf_b200.engine import Engine
    sz, K, T = [24, 16, 5], 5, 6
    pos, sig, beta, C, frames = _case(sz, K, T, 61)
    data = tmp_path / "data.npz"
    np.savez(data, sz=np.asarray(sz), K=K, T=T, pos=pos, sig=sig, beta=beta.numpy(), C=C.numpy(), frames=frames.numpy())
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(worker), ROOT, str(data),
                          str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    e.upload_frames(frames, clamp_negative=False)
    b, c = beta.clone().cuda(), C.cuda()
    m, v = torch.zeros_like(b), torch.zeros_like(b)
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    ids = torch.arange(T, dtype=torch.int32, device="cuda")
    ref_losses = []
    for step in range(1, 4):
        e.motion_step(ids, b, m, v, c, 1e-3, (0.9, 0.999), 1e-8, step, False, B_global=T, loss_out=loss)
        ref_losses.append(float(loss))
    parts = [np.load(str(tmp_path / "out") + "_%d.npz" % r) for r in range(2)]
    got = np.concatenate([p["beta"] for p in parts], 2)
    assert np.array_equal(got, b.cpu().numpy())
    np.testing.assert_allclose(parts[0]["losses"], ref_losses, rtol=1e-12)


def test_binning_large_k_bit_exact():
    """config-4-like density: K=1000 wide footprints on a 64x32x21 sub-volume."""
    from dnmf_b200.engine import Engine
    rng = np.random.default_rng(4)
    sz, K, T = [64, 32, 21], 1000, 2
    pos = (rng.random((K, 3)) * np.asarray(sz) * 1.2 - 0.1 * np.asarray(sz)).astype(np.float32)
    sig = np.full(K, 6.0, np.float32)
    e = Engine(sz, K, T)
    e.set_footprints(pos, sig, 3.5)
    _, _, beta, _, _ = _case(sz, 2, T, 77, beta_scale=0.5)
    counts, offsets, ids, wins = e.bin_tiles(beta.cuda(), torch.arange(T))
    tl = e.tiling()
    rc, ro, ri, rw = O.bin_tiles(beta.numpy(), [0, 1], e.ranges(), sz, (tl["tx"], tl["ty"], tl["tz"]))
    assert np.array_equal(wins, rw) and np.array_equal(counts, rc)
    assert np.array_equal(offsets, ro) and np.array_equal(ids, ri)
    assert counts.max() > 200


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0), (2, 4, 0, 0), (1, 1, 0, 0, 2)])
def test_dense_neurons_loss_grad(tiling):
    """many overlapping wide footprints (long lists, staged + overflow slots) against the closed form."""
    _check([40, 24, 6], 120, 2, seed=13, cutoff=3.5, tiling=tiling, sigma=5.0, tol=5e-5)


@pytest.mark.parametrize("sz", [[256, 128, 21], [512, 256, 32], [50, 50, 2], [33, 17, 9]])
def test_exact_fast_division_is_proven_for_config_sizes(sz):
    """the 3-instruction division + folded multiply is enabled only after the device proved, over all 2^32
    inputs per axis, that it reproduces the reference op sequence bit for bit."""
    from dnmf_b200.engine import Engine
    e = Engine(sz, 2, 1)
    e.set_footprints(np.array([[1., 1., 1.], [2., 2., 1.]], np.float32), np.ones(2, np.float32), 3.5)
    assert e.tiling()["fast_div"] == 1


def test_launch_chunking_over_grid_z_limit():
    """B * ntz > 65535 forces the fused kernel out in several launches; per-frame results must equal the
    ones computed in small batches (fixed-order reductions make them bit-identical)."""
    from dnmf_b200.engine import Engine
    sz, K, T = [16, 8, 32], 3, 2100
    pos, sig, _, _, _ = _case(sz, K, 1, 71)
    g = torch.Generator().manual_seed(71)
    frames = torch.rand(T, *sz, generator=g)
    s = torch.tensor([.3, .01, .01, .01, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4])[:, None, None]
    beta = (O.identity_beta(T) + s * torch.randn(10, 3, T, generator=g)).cuda()
    C = torch.rand(K, T, generator=g).cuda()
    e = Engine(sz, K, T)
    e.set_tiling(1, 1, 1, 0, 1)                       # tz = 1 -> ntz = 32 -> 67200 grid-z slices
    e.set_footprints(pos, sig, 3.5)
    e.upload_frames(frames, clamp_negative=False)
    g_all, s_all = e.loss_grad(torch.arange(T), beta, C)
    for lo in (0, 1000, 2000):
        ids = torch.arange(lo, lo + 100)
        g_part, s_part = e.loss_grad(ids, beta, C, B_global=T)
        assert torch.equal(s_part, s_all[lo:lo + 100])
        assert torch.equal(g_part[:, :, lo:lo + 100], g_all[:, :, lo:lo + 100])


def test_frames_per_cta_do_not_change_results(monkeypatch):
    """One CTA walking 1, 3 or 8 consecutive frames of its tile (slices cached across frames, traces and tile
    requested a frame ahead) gives bit-identical gradients."""
    from dnmf_b200.engine import Engine
    sz, K, T = [48, 24, 6], 9, 16
    pos, sig, beta, C, frames = _case(sz, K, T, 77, beta_scale=0.3)
    out = []
    for fpc in ("1", "3", "8"):
        monkeypatch.setenv("DNMF_FPC", fpc)
        e = Engine(sz, K, T)
        e.set_tiling(1, 1, 0, 0, 2)
        e.set_footprints(pos, sig, 3.5)
        out.append(e.loss_grad(torch.arange(T), beta.cuda(), C.cuda(), frames=frames.cuda()))
    for g, s_ in out[1:]:
        assert torch.equal(g, out[0][0]) and torch.equal(s_, out[0][1])


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0), (1, 1, 0, 0, 2), (2, 2, 0, 0, 2)])
def test_depth_32_padded_tile_and_rotated_z_order(tiling):
    """Z = 32 makes the dense shared-memory pitches multiples of 32 floats: the tile is stored with a padded x
    pitch and the lanes walk z in rotated order (no tensor-map copy); Z = 64 adds depth-chunked tiles."""
    _check([24, 16, 32], 6, 3, seed=32, cutoff=3.5, tiling=tiling, beta_scale=0.5)
    _check([16, 8, 64], 5, 2, seed=64, cutoff=3.0, tiling=tiling)


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0, 2), (2, 2, 0, 0, 2), (1, 1, 0, 3, 2)])
def test_cached_slices_across_frames_match_oracle(monkeypatch, tiling):
    """Runs of frames with the same deformation but different traces: the CTA keeps the staged slices and the
    list and only refreshes the traces (the hot path of a near-converged fit).  Mixed with frames whose window or
    list changes, and with a slot capacity of 3 that sends part of the lists through the overflow path.
    Checked against the oracle and against one-frame-per-CTA launches."""
    from dnmf_b200.engine import Engine
    sz, K, T = [40, 24, 9], 10, 12
    pos, sig, beta, C, frames = _case(sz, K, T, 91, sigma=2.5, beta_scale=0.6)
    beta[:, :, 1] = beta[:, :, 0]          # identity run: frames 0, 1
    beta[:, :, 4] = beta[:, :, 3]          # a deformed run: frames 3, 4, 5
    beta[:, :, 5] = beta[:, :, 3]
    beta[:, :, 9] = beta[:, :, 8]
    beta[0, 0, 9] += 1e-3                  # almost the same: same window, different samples
    tabs, _ = O.axis_tables(pos, sig, sz, 3.5)
    loss, gref = O.closed_form_step(frames.numpy(), list(range(T)), beta.numpy(), C.numpy(), tabs, sz)
    out = []
    for fpc in ("1", "4", "12"):
        monkeypatch.setenv("DNMF_FPC", fpc)
        e = Engine(sz, K, T)
        e.set_tiling(*tiling)
        e.set_footprints(pos, sig, 3.5)
        grad, sse = e.loss_grad(torch.arange(T), beta.cuda(), C.cuda(), frames=frames.cuda())
        N = int(np.prod(sz))
        assert abs(float(sse.sum()) / (T * N) - loss) <= 1e-5 * loss
        err = np.abs(grad.cpu().numpy() - gref).max() / np.abs(gref).max()
        assert err < 3e-5, err
        out.append((grad, sse))
    for g, s_ in out[1:]:
        assert torch.equal(g, out[0][0]) and torch.equal(s_, out[0][1])


@pytest.mark.parametrize("case", ["full", "ragged_subset", "affine", "repeated_frame", "single_batch"])
def test_epoch_call_equals_per_step_calls(case):
    """dnmf_motion_epoch (all minibatches of an epoch in one library call, resident video) leaves beta, the Adam
    moments and the per-step losses bit-identical to one dnmf_motion_step per minibatch -- batch by batch and
    frame-parallel (one fused launch over the epoch, each column replaying the zero-gradient Adam steps of the
    other minibatches).  Two epochs, so that the second starts from non-zero moments and step numbers; ragged last
    batch, frames left out of the epoch, frozen quadratic rows, and an epoch that draws a frame twice (must fall
    back to batch by batch)."""
    from dnmf_b200.engine import Engine
    sz, K, T, B = [32, 16, 5], 6, 14, 4
    pos, sig, beta0, C, frames = _case(sz, K, T, 5, beta_scale=0.2)
    gen = torch.Generator().manual_seed(1)
    epochs = []
    for ep in range(2):
        ids = torch.randperm(T, generator=gen).to(torch.int32)
        if case == "full":
            ids, off = ids[:12], list(range(0, 13, B))
        elif case == "ragged_subset":
            ids, off = ids[:11], [0, 4, 8, 11]
        elif case == "affine":
            off = [0, 4, 8, 12, 14]
        elif case == "repeated_frame":
            ids = torch.cat([ids[:8], ids[2:6]])
            off = [0, 4, 8, 12]
        else:
            ids, off = ids[:4], [0, 4]
        epochs.append((ids, off))
    affine = case == "affine"
    res = []
    for mode in ("steps", "epoch_sequential", "epoch"):
        e = Engine(sz, K, T)
        e.set_footprints(pos, sig, 3.5)
        e.upload_frames(frames, clamp_negative=False)
        beta = beta0.clone().cuda()
        m, v = torch.zeros_like(beta), torch.zeros_like(beta)
        c = C.cuda()
        all_losses = []
        step = 1
        for ids, off in epochs:
            nb = len(off) - 1
            losses = torch.zeros(nb, dtype=torch.float64, device="cuda")
            idd = ids.cuda()
            if mode == "steps":
                for i in range(nb):
                    e.motion_step(idd[off[i]:off[i + 1]], beta, m, v, c, 1e-3, (0.9, 0.999), 1e-8, step + i,
                                  affine=affine, loss_out=losses[i:i + 1])
            else:
                e.epoch_mode(1 if mode == "epoch_sequential" else 0)
                e.motion_epoch(idd, off, beta, m, v, c, 1e-3, (0.9, 0.999), 1e-8, step, affine=affine,
                               loss_out=losses)
                expect_parallel = mode == "epoch" and case not in ("repeated_frame", "single_batch")
                assert e.epoch_mode() == (1 if expect_parallel else 0)
            step += nb
            all_losses.append(losses)
        res.append((beta, m, v, torch.cat(all_losses)))
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert torch.equal(a, b)
    assert float(res[0][3].min()) > 0
    if affine:
        assert torch.equal(res[0][0][4:].cpu(), beta0[4:])


def _mu_both_paths(sz, K, T, seed, tiling, sigma=2.5, cutoff=3.5, beta_scale=1.0, expect_fused=True, resident=False):
    """Trace statistics through the fused-tile kernel and through the panel kernel, both against the oracle."""
    from dnmf_b200.engine import Engine
    pos, sig, beta, C, frames = _case(sz, K, T, seed, sigma, beta_scale)
    e = Engine(sz, K, T)
    e.set_tiling(*tiling)
    e.set_footprints(pos, sig, cutoff)
    ids = torch.arange(T)
    dev_frames = frames.cuda()
    if resident:
        e.upload_frames(frames, clamp_negative=False)
        dev_frames = None
    tabs, _ = O.axis_tables(pos, sig, sz, cutoff)
    Gm, bv = O.closed_form_mu_stats(frames.numpy(), list(range(T)), beta.numpy(), tabs, sz)
    res = []
    for force_panel in (0, 1):
        e.mu_path(force_panel)
        e.mu_stats(ids, beta.cuda(), frames=dev_frames)
        path = e.mu_path() & 1
        if force_panel == 0 and e.tiling()["fast_div"]:
            assert expect_fused is None or path == (1 if expect_fused else 0), (path, e.tiling())
        if force_panel == 1:
            assert path == 0
        G = np.stack([e.get_mu_stats(t)[0] for t in range(T)])
        b = np.stack([e.get_mu_stats(t)[1] for t in range(T)])
        np.testing.assert_allclose(G, Gm, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(b, bv, rtol=2e-5, atol=1e-6)
        assert np.array_equal(G, np.swapaxes(G, 1, 2)) or np.allclose(G, np.swapaxes(G, 1, 2), rtol=1e-6, atol=1e-9)
        res.append((G, b))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=2e-5, atol=1e-6)
    return e


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0, 2), (2, 1, 0, 0, 2), (2, 2, 0, 0, 2)])
def test_mu_stats_fused_tiles_short_lists(tiling):
    """fit_tile_kernel<MODE=3>: one diagonal block of 1, 2 or 3 slot pairs (lists of 1..6 neurons), ragged
    volumes, a batch passed by pointer and the resident video."""
    _mu_both_paths([33, 18, 5], 4, 3, seed=7, tiling=tiling)
    _mu_both_paths([50, 50, 2], 10, 4, seed=8, tiling=tiling, sigma=3.0, resident=True,
                   beta_scale=1.0 if tiling[1] == 1 else 0.2)
    _mu_both_paths([17, 30, 11], 2, 2, seed=9, tiling=tiling)


@pytest.mark.parametrize("tiling", [(1, 1, 0, 16, 2), (2, 2, 0, 16, 2)])
def test_mu_stats_fused_tiles_long_lists(tiling):
    """lists of 7..16 neurons: several row blocks and the mirrored off-diagonal blocks."""
    e = _mu_both_paths([32, 24, 6], 14, 3, seed=21, tiling=tiling, sigma=4.0, cutoff=0.0)
    assert e.tiling()["cap"] >= 14
    _mu_both_paths([40, 24, 9], 30, 2, seed=22, tiling=tiling, sigma=3.0, beta_scale=0.5)


def test_mu_stats_fused_tiles_capacity_redo_and_depth_32():
    """The statistics launch stages every listed neuron (its own capacity, larger than the fit's: 40 slots where
    the fit was given 4).  A window wider than the staged slices raises the overflow flag and the panel kernel
    redoes the call.  Z = 32 runs the rotated z order of the padded tile."""
    _mu_both_paths([24, 16, 6], 40, 2, seed=9, tiling=(1, 1, 0, 4, 2), sigma=4.0, cutoff=0.0)
    _mu_both_paths([24, 16, 32], 6, 3, seed=32, tiling=(1, 1, 0, 0, 2), beta_scale=0.5)
    _mu_both_paths([24, 16, 6], 8, 3, seed=11, tiling=(1, 1, 0, 0, 2), beta_scale=8.0, expect_fused=False)


@pytest.mark.parametrize("gamma", [None, 1e-2])
def test_mu_sweeps_sparse_neighbour_lists_match_dense_and_oracle(gamma):
    """The sweeps over G compacted to the static neighbour lists (neurons whose truncated supports overlap) equal
    the dense sweeps and the oracle's multiplicative update (Demix/dNMF.py:143-148), with and without the temporal
    coupling; without a cutoff every neuron neighbours every other and the dense kernel is kept."""
    from dnmf_b200.engine import Engine
    sz, K, T, iters = [48, 32, 6], 40, 5, 7
    pos, sig, beta, C, frames = _case(sz, K, T, 77, sigma=2.0, beta_scale=0.5)
    pos[-1] = [-40.0, -40.0, -40.0]                    # a neuron outside the volume: empty range, no neighbours
    for cutoff, expect_sparse in ((3.0, True), (0.0, False)):
        e = Engine(sz, K, T)
        e.set_tiling(1, 1, 0, 0, 2)
        e.set_footprints(pos, sig, cutoff)
        e.mu_stats(torch.arange(T), beta.cuda(), frames=frames.cuda())
        G = np.stack([e.get_mu_stats(t)[0] for t in range(T)])
        b = np.stack([e.get_mu_stats(t)[1] for t in range(T)])
        ref = C.numpy().astype(np.float64)
        for _ in range(iters):
            ref = O.mu_sweep(np.transpose(G, (1, 2, 0)), b.T, ref, gamma)
        out = []
        for flags in (0, 4, 2):   # automatic (all sweeps of a frame in one CTA when uncoupled), per-sweep launches, dense
            e.mu_path(flags)
            c = C.clone().cuda()
            e.mu_sweeps(c, gamma, iters)
            assert ((e.mu_path() >> 1) & 1) == (1 if (flags != 2 and expect_sparse) else 0)
            out.append(c.cpu().numpy())
            np.testing.assert_allclose(out[-1], ref.astype(np.float32), rtol=2e-6, atol=1e-30)
        np.testing.assert_allclose(out[0], out[1], rtol=1e-6, atol=1e-30)
        np.testing.assert_allclose(out[0], out[2], rtol=1e-6, atol=1e-30)
        if expect_sparse:                           # gamma = 0.0 is the uncoupled update too (demo.py:46)
            e.mu_path(0)
            c0 = C.clone().cuda()
            e.mu_sweeps(c0, 0.0, iters)
            cn = C.clone().cuda()
            e.mu_sweeps(cn, None, iters)
            assert torch.equal(c0, cn)


@pytest.mark.parametrize("K", [40, 140])
def test_mu_stats_panel_kernel_8x8_blocks_long_lists(monkeypatch, K):
    """Panel kernel with 8x8 register blocks over the [voxel][row] panel (lists of 32 rows and more): split-K
    over the voxels when there are few blocks (K = 40), two blocks per thread (K = 140); against the oracle and
    against the 4x4 variant."""
    from dnmf_b200.engine import Engine
    sz, T = [24, 16, 6], 2
    pos, sig, beta, C, frames = _case(sz, K, T, 9, sigma=4.0)
    tabs, _ = O.axis_tables(pos, sig, sz, 0.0)
    Gm, bv = O.closed_form_mu_stats(frames.numpy(), list(range(T)), beta.numpy(), tabs, sz)
    out = []
    for block4 in ("0", "1"):
        monkeypatch.setenv("DNMF_MU_BLOCK4", block4)
        e = Engine(sz, K, T)
        e.set_tiling(1, 1, 0, 0, 2)
        e.set_footprints(pos, sig, 0.0)
        e.mu_path(1)                                    # panel kernel
        e.mu_stats(torch.arange(T), beta.cuda(), frames=frames.cuda())
        assert e.mu_path() & 1 == 0
        G = np.stack([e.get_mu_stats(t)[0] for t in range(T)])
        b = np.stack([e.get_mu_stats(t)[1] for t in range(T)])
        np.testing.assert_allclose(G, Gm, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(b, bv, rtol=2e-5, atol=1e-6)
        out.append(G)
    np.testing.assert_allclose(out[0], out[1], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0, 2), (2, 2, 0, 0, 2), (1, 1, 0, 3, 2)])
def test_main_loop_bodies_every_tail_kind(tiling):
    """Lists of every tail kind (slot capacity 3: even and odd lists, single slots, overflow through the global tables)
    through the specialised main-loop bodies: against the closed form, and bit-identical over repeated launches."""
    from dnmf_b200.engine import Engine
    sz, K, T = [40, 24, 9], 10, 12
    pos, sig, beta, C, frames = _case(sz, K, T, 17, sigma=2.5, beta_scale=0.6)
    tabs, _ = O.axis_tables(pos, sig, sz, 3.5)
    loss, gref = O.closed_form_step(frames.numpy(), list(range(T)), beta.numpy(), C.numpy(), tabs, sz)
    e = Engine(sz, K, T)
    e.set_tiling(*tiling)
    e.set_footprints(pos, sig, 3.5)
    out = [e.loss_grad(torch.arange(T), beta.cuda(), C.cuda(), frames=frames.cuda()) for _ in range(4)]
    torch.cuda.synchronize()
    for g, s_ in out[1:]:
        assert torch.equal(g, out[0][0]) and torch.equal(s_, out[0][1])
    err = np.abs(out[0][0].cpu().numpy() - gref).max() / np.abs(gref).max()
    assert err < 3e-5, err
