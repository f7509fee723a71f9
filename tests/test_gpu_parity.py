"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle and the committed goldens.

Tolerances (BASELINE.json north_star): loss per iteration <= 1e-4 relative, deformation field
<= 1e-3 px, traces <= 1e-3 relative.  Integer work (ranges, windows, bin lists) is bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import dnmf_oracle as O

pytestmark = pytest.mark.gpu


def _engine(sz, K, T, pos, sigma, cutoff, tiling=(1, 1, 0, 0)):
    from dnmf_b200.engine import Engine
    e = Engine(sz, K, T)
    e.set_tiling(*tiling)
    e.set_footprints(pos, sigma, cutoff)
    return e


def _rand_beta(T, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    s = torch.tensor([1.0, .02, .02, .02, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4, 5e-4])[:, None, None] * scale
    return O.identity_beta(T) + s * torch.randn(10, 3, T, generator=g)


@pytest.mark.parametrize("cutoff", [0.0, 3.5, 2.0])
def test_tables_and_ranges(golden_random, cutoff):
    g = golden_random
    sz = g["sz"].tolist()
    e = _engine(sz, 5, 6, g["pos"], g["sigma"], cutoff)
    tabs, rng = O.axis_tables(g["pos"], g["sigma"], sz, cutoff)
    assert np.array_equal(e.ranges(), rng)          # integer: bit-exact
    for d in range(3):
        np.testing.assert_allclose(e.table(d), tabs[d], rtol=3e-6, atol=1e-12)


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0), (2, 1, 0, 0), (2, 2, 4, 0), (2, 4, 0, 0), (1, 1, 0, 0, 2)])
def test_binning_bit_exact(golden_random, tiling):
    g = golden_random
    sz = g["sz"].tolist()
    e = _engine(sz, 5, 6, g["pos"], g["sigma"], 2.0, tiling)
    beta = torch.tensor(g["beta"]).cuda()
    times = [0, 3, 5, 1]
    counts, offsets, ids, wins = e.bin_tiles(beta, torch.tensor(times))
    tl = e.tiling()
    rc, ro, ri, rw = O.bin_tiles(g["beta"], times, e.ranges(), sz, (tl["tx"], tl["ty"], tl["tz"]))
    assert np.array_equal(wins, rw)
    assert np.array_equal(counts, rc)
    assert np.array_equal(offsets, ro)
    assert np.array_equal(ids, ri)


@pytest.mark.parametrize("tiling", [(1, 1, 0, 0), (2, 1, 3, 0), (2, 2, 0, 1), (2, 4, 0, 0), (1, 1, 0, 0, 2), (1, 1, 4, 2, 2)])
def test_loss_grad_vs_reference_autograd(golden_random, tiling):
    """random quadratic beta incl. 42 % out-of-bounds samples and one identity frame (F2)."""
    g = golden_random
    sz = g["sz"].tolist()
    T = g["beta"].shape[2]
    e = _engine(sz, 5, T, g["pos"], g["sigma"], 0.0, tiling)
    beta = torch.tensor(g["beta"]).cuda()
    C = torch.tensor(g["C"]).cuda()
    frames = torch.tensor(g["frames"]).cuda()
    ids = torch.arange(T)
    grad, sse = e.loss_grad(ids, beta, C, frames=frames)
    N = int(np.prod(sz))
    loss = float(sse.sum()) / (T * N)
    assert abs(loss - float(g["loss"])) <= 1e-5 * float(g["loss"])
    ref = g["grad"]
    err = np.abs(grad.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err
    AtC, At, grid = e.forward(ids, beta, C, want_At=True, want_grid=True)
    np.testing.assert_allclose(AtC.cpu().numpy(), g["A_tC"], atol=2e-6)
    np.testing.assert_allclose(At.cpu().numpy(), g["A_t"], atol=2e-6)


def test_identity_floor_cell_selection(golden_demo):
    """at identity init the gradient depends on fp32 round-off of the coordinate
    pipeline; compare with the closed-form oracle that replays the op order (itself pinned to
    the real reference by tests/test_oracle.py)."""
    g = golden_demo
    sz = g["sz"].tolist()
    K, T = g["C0"].shape
    sigma = np.full(K, 3.0, np.float32)
    e = _engine(sz, K, T, g["pos0"], sigma, 0.0)
    beta = O.identity_beta(T).cuda()
    C = torch.tensor(g["C0"]).cuda()
    e.upload_frames(torch.tensor(g["frames"]))
    times = [0, 1, 2, 3]
    grad, sse = e.loss_grad(torch.tensor(times), beta, C)
    tabs, _ = O.axis_tables(g["pos0"], sigma, sz, 0.0)
    l, gr = O.closed_form_step(g["frames"][times], times, O.identity_beta(T).numpy(), g["C0"], tabs, sz)
    N = int(np.prod(sz))
    assert abs(float(sse.sum()) / (4 * N) - l) <= 1e-6 * l
    assert abs(l - g["losses"][0]) <= 1e-6 * l
    err = np.abs(grad.cpu().numpy() - gr).max() / np.abs(gr).max()
    assert err < 5e-6, err


def test_adam_matches_torch():
    from dnmf_b200.engine import Engine
    T = 7
    e = Engine([8, 8, 2], 2, T)
    torch.manual_seed(0)
    p = torch.randn(10, 3, T)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    dp = p.clone().cuda()
    m = torch.zeros_like(dp)
    v = torch.zeros_like(dp)
    for step in range(1, 6):
        g = torch.randn(10, 3, T) * (10.0 ** -step)
        g[:, :, 3] = 0                      # off-batch frame: g = 0 but momentum still moves it (F4)
        ref_p.grad = g.clone()
        opt.step()
        dg = g.clone().cuda()
        e.adam_step(dp, dg, m, v, 1e-3, (0.9, 0.999), 1e-8, step)
        assert float(dg.abs().max()) == 0.0   # consumed and reset
        np.testing.assert_allclose(dp.cpu().numpy(), ref_p.detach().numpy(), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(m.cpu().numpy(), opt.state[ref_p]["exp_avg"].numpy(), rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(v.cpu().numpy(), opt.state[ref_p]["exp_avg_sq"].numpy(), rtol=2e-6, atol=1e-14)


def _demo_model(g, cutoff, resident, tiling=None):
    from dnmf_b200 import DeformableNMF
    sz = g["sz"].tolist()
    K, T = g["C0"].shape
    dn = DeformableNMF(sz, K, T, positions=torch.tensor(g["pos0"]), cutoff=cutoff, verbose=False, tiling=tiling)
    dn.C = torch.tensor(g["C0"]).cuda()
    if resident:
        dn.attach_video(torch.tensor(g["frames"]), layout="TXYZ")
    return dn


@pytest.mark.parametrize("cutoff,resident", [(0.0, False), (3.5, True), (3.5, False)])
def test_demo_trajectory_vs_reference(golden_demo, cutoff, resident):
    """demo.py:41-46 schedule, first 250 Adam steps (shuffle=False, batch 4, lr 1e-5) then the trace
    updates, against the REAL reference's recorded outputs (tests/golden/demo_cfg1.npz)."""
    from torch.utils.data import DataLoader
    from dnmf_b200 import FrameDataset
    g = golden_demo
    dn = _demo_model(g, cutoff, resident)
    loader = DataLoader(FrameDataset(torch.tensor(g["frames"])), batch_size=4, shuffle=False)
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    dn.update_motion(loader, opt, gamma=1, epochs=10)
    losses = dn.losses()
    ref = g["losses"]
    assert losses.shape == ref.shape
    rel = np.abs(losses - ref) / ref
    assert rel.max() < 1e-4, rel.max()                     # north_star: loss per iteration <= 1e-4 relative
    beta = dn.fp.beta.detach().cpu()
    # deformation field tau_t(p) over all voxels and frames, <= 1e-3 px
    _, phi = O.voxel_basis(g["sz"].tolist())
    q_new = torch.einsum("mnza,abt->mnzbt", phi, beta)
    q_ref = torch.einsum("mnza,abt->mnzbt", phi, torch.tensor(g["beta250"]))
    assert float((q_new - q_ref).abs().max()) < 1e-3
    st = opt.state[dn.fp.beta]
    assert int(st["step"]) == 250
    m = st["exp_avg"].cpu().numpy()
    assert np.abs(m - g["adam_m"]).max() <= 1e-3 * np.abs(g["adam_m"]).max()
    # traces: update_footprints(gamma_c=0, iter_c=50) then (gamma_c=1e-2, iter_c=10)
    dn.update_footprints(loader, 4, g["sz"].tolist(), gamma_c=0, iter_c=50, dense=False)
    C = dn.C.cpu().numpy()
    assert np.abs(C - g["C_mu0"]).max() / np.abs(g["C_mu0"]).max() < 1e-3
    dn.update_footprints(loader, 4, g["sz"].tolist(), gamma_c=1e-2, iter_c=10, dense=False)
    C = dn.C.cpu().numpy()
    assert np.abs(C - g["C_mu1"]).max() / np.abs(g["C_mu1"]).max() < 1e-3


def test_mu_stats_vs_closed_form(golden_random):
    g = golden_random
    sz = g["sz"].tolist()
    T = g["beta"].shape[2]
    e = _engine(sz, 5, T, g["pos"], g["sigma"], 0.0)
    beta = torch.tensor(g["beta"]).cuda()
    frames = torch.tensor(g["frames"]).cuda()
    e.mu_stats(torch.arange(T), beta, frames=frames)
    tabs, _ = O.axis_tables(g["pos"], g["sigma"], sz, 0.0)
    Gm, bv = O.closed_form_mu_stats(g["frames"], list(range(T)), g["beta"], tabs, sz)
    for t in range(T):
        G, b = e.get_mu_stats(t)
        np.testing.assert_allclose(G, Gm[t], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(b, bv[t], rtol=2e-5, atol=1e-6)
    # and against the reference's dense A_t
    A = g["A_t"].astype(np.float64).reshape(T, 5, -1)
    for t in range(T):
        G, b = e.get_mu_stats(t)
        np.testing.assert_allclose(G, A[t] @ A[t].T, rtol=1e-4, atol=1e-5)


def test_update_footprints_dense_outputs(golden_demo):
    """A_t / Y returned by update_footprints match the reference's dense arrays (first 4 frames)."""
    from torch.utils.data import DataLoader
    from dnmf_b200 import FrameDataset
    g = golden_demo
    dn = _demo_model(g, 0.0, False)
    dn.fp.beta.data.copy_(torch.tensor(g["beta250"]))
    frames = torch.tensor(g["frames"][:8])
    loader = DataLoader(FrameDataset(frames), batch_size=4, shuffle=False)
    A_t, Y_i, Y = dn.update_footprints(loader, 4, g["sz"].tolist(), gamma_c=0, iter_c=1)
    assert A_t.shape == (50, 50, 2, 10, 8) and Y.shape == (50, 50, 2, 8) and Y_i.shape == Y.shape
    np.testing.assert_allclose(A_t[..., :4], g["A_t_first4"], atol=3e-6)
    np.testing.assert_allclose(Y[..., :4], g["Y_first4"], atol=0)


def test_shuffled_loader_matches_reference(golden_demo):
    """demo.py:34 (shuffle=True): replay the reference's recorded batch order -- and check that a torch
    DataLoader seeded the same way produces that order -- for two epochs of random minibatches."""
    import os
    from torch.utils.data import DataLoader
    from dnmf_b200 import FrameDataset
    g = golden_demo
    gs = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "demo_shuffle.npz")))
    dn = _demo_model(g, 3.5, False)
    ds = FrameDataset(torch.tensor(g["frames"]))
    loader = DataLoader(ds, batch_size=4, shuffle=True, num_workers=0)
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    order = []

    class Recorder:                                   # same loader, re-seeded per epoch like the golden run
        def __init__(self, ep):
            self.ep = ep

        def __iter__(self):
            torch.manual_seed(1000 + self.ep)
            for frames, idx in loader:
                order.append(idx.numpy().copy())
                yield frames, idx

        def __len__(self):
            return len(loader)

    for ep in range(2):
        dn.update_motion(Recorder(ep), opt, gamma=1, epochs=1)
    assert np.array_equal(np.asarray(order), gs["order"])
    losses = dn.losses()
    assert np.max(np.abs(losses - gs["losses"]) / gs["losses"]) < 1e-4
    assert float((dn.fp.beta.detach().cpu() - torch.tensor(gs["beta"])).abs().max()) < 1e-6


def test_opt_in_jacobian_regularizer_vs_autograd(golden_demo):
    """OPT-IN (off by default, not reference behaviour): differentiable index-consistent log-det-Jacobian
    penalty.  With the flag ON and gamma > 0 the updates follow the oracle's autograd version; with the flag
    OFF gamma is inert like in the reference."""
    from torch.utils.data import DataLoader
    from dnmf_b200 import DeformableNMF, FrameDataset
    g = golden_demo
    sz = g["sz"].tolist()
    K, T = g["C0"].shape
    frames = torch.tensor(g["frames"][:16])
    gamma = 1e-3
    port = O.TorchPort(sz, K, 16, positions=g["pos0"], C0=g["C0"][:, :16])
    popt = torch.optim.Adam([port.beta], lr=1e-4)
    ref_losses = [port.motion_step(frames[i:i + 4], list(range(i, i + 4)), popt, reg_gamma=gamma)
                  for _ in range(3) for i in range(0, 16, 4)]
    dn = DeformableNMF(sz, K, 16, positions=torch.tensor(g["pos0"]), cutoff=0.0, verbose=False,
                       jacobian_regularizer=True)
    dn.C = torch.tensor(g["C0"][:, :16]).contiguous().cuda()
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-4)
    dn.update_motion(DataLoader(FrameDataset(frames), batch_size=4, shuffle=False), opt, gamma=gamma, epochs=3)
    assert np.max(np.abs(dn.losses() - np.asarray(ref_losses)) / np.asarray(ref_losses)) < 1e-4
    assert float((dn.fp.beta.detach().cpu() - port.beta.detach()).abs().max()) < 2e-6
    # flag OFF: gamma changes nothing
    a = DeformableNMF(sz, K, 16, positions=torch.tensor(g["pos0"]), cutoff=0.0, verbose=False)
    b = DeformableNMF(sz, K, 16, positions=torch.tensor(g["pos0"]), cutoff=0.0, verbose=False)
    for m, gm in ((a, 0), (b, 5.0)):
        m.C = torch.tensor(g["C0"][:, :16]).contiguous().cuda()
        m.update_motion(DataLoader(FrameDataset(frames), batch_size=4, shuffle=False),
                        torch.optim.Adam([m.fp.beta], lr=1e-4), gamma=gm, epochs=1)
    assert torch.equal(a.fp.beta, b.fp.beta)


def test_reg_return_of_forward_matches_reference(golden_random, golden_demo, golden_extras):
    """`reg` of ExponentialFP.forward (Demix/dNMF.py:60-61: log-det-Jacobian at two corners with the reference's
    own cross-term indexing) against the real reference's output, for a random quadratic beta and for the demo's
    beta after 250 Adam steps."""
    from dnmf_b200 import ExponentialFP
    g, d, x = golden_random, golden_demo, golden_extras
    fp = ExponentialFP(g["sz"].tolist(), 5, 6, positions=torch.tensor(g["pos"]), shape_std=2.5, cutoff=0.0)
    with torch.no_grad():
        fp.beta.copy_(torch.tensor(g["beta"]).cuda())
    _, _, _, reg = fp(list(range(6)), torch.tensor(g["C"]))
    assert not reg.is_cuda and not reg.requires_grad          # detached, on the CPU, like the reference's (F3)
    np.testing.assert_allclose(reg.numpy(), x["reg_random"], rtol=1e-5, atol=1e-9)
    fpd = ExponentialFP(d["sz"].tolist(), 10, 100, positions=torch.tensor(d["pos0"]))
    with torch.no_grad():
        fpd.beta.copy_(torch.tensor(d["beta250"]).cuda())
    _, _, _, regd = fpd(list(range(8)), torch.tensor(d["C0"]))
    np.testing.assert_allclose(regd.numpy(), x["reg_demo250_first8"], rtol=1e-4, atol=1e-9)


def test_registered_video_matches_reference_nearest_neighbour(golden_random, golden_extras):
    """Y_i of spatial_pushforward (Demix/dNMF.py:81-83,90-103: scipy NearestNDInterpolator over the deformed
    points) against the real reference on a 3-D volume whose deformed points are in general position.  Exact
    equality is required wherever the nearest neighbour is unique; the voxels with a distance tie (counted with
    a float64 brute force here) may differ, there the reference's KD-tree order is arbitrary."""
    from dnmf_b200.engine import Engine
    g, x = golden_random, golden_extras
    sz = g["sz"].tolist()
    X, Y, Z = sz
    T = x["pf_beta"].shape[2]
    e = Engine(sz, 5, T)
    e.set_footprints(g["pos"], g["sigma"], 0.0)
    frames = torch.tensor(g["frames"])
    yi = e.iwarp(torch.arange(T), torch.tensor(x["pf_beta"]).cuda(), frames=frames.cuda()).cpu().numpy()
    ref = np.moveaxis(x["pf_Y_i"], 3, 0)
    grid = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    _, phi = O.voxel_basis(sz)
    ties = mism = mism_unique = 0
    for t in range(T):
        q = torch.einsum("mnza,ab->mnzb", phi, torch.tensor(x["pf_beta"][:, :, t]))
        u = 2 * q / (torch.tensor(sz)[None, None, None, :] - 1) - 1
        P = (((u + 1) / 2) * torch.tensor(sz)[None, None, None, :].float()).reshape(-1, 3).numpy().astype(np.float64)
        d2 = ((grid[:, None, :] - P[None, :, :]) ** 2).sum(-1)
        tie = ((d2 <= d2.min(1)[:, None]).sum(1) > 1).reshape(X, Y, Z)
        bad = yi[t] != ref[t]
        ties += int(tie.sum())
        mism += int(bad.sum())
        mism_unique += int((bad & ~tie).sum())
    print("\n[Y_i] %d voxels, %d with a distance tie, %d differ from the reference, %d of them where the nearest "
          "neighbour is unique" % (T * X * Y * Z, ties, mism, mism_unique))
    assert mism_unique == 0
    assert mism <= ties


def test_demo_registered_video_first_plane_matches_reference(golden_demo):
    """Y_i at the demo size (50x50x2): the reference scales the deformed points by sz, not sz-1
    (Demix/dNMF.py:83), which puts the z = 1 sources at z = 2: every z = 1 voxel is equidistant from both planes
    (a tie, KD-tree order arbitrary), the z = 0 plane has unique neighbours and must match exactly."""
    g = golden_demo
    dn = _demo_model(g, 3.5, resident=False)
    # the golden's Y_i was produced after the 250 Adam steps: use the reference's own beta for the comparison
    with torch.no_grad():
        dn.fp.beta.copy_(torch.tensor(g["beta250"]).cuda())
    yi = dn.fp.engine.iwarp(torch.arange(4), dn.fp.beta.detach(), frames=torch.tensor(g["frames"][:4]).cuda()).cpu().numpy()
    ref = np.moveaxis(g["Y_i_first4"], 3, 0)
    same0 = float((yi[:, :, :, 0] == ref[:, :, :, 0]).mean())
    same1 = float((yi[:, :, :, 1] == ref[:, :, :, 1]).mean())
    print("\n[Y_i demo] plane z=0 equal %.4f, plane z=1 (all ties) equal %.4f" % (same0, same1))
    assert same0 == 1.0


def test_traces_from_the_references_own_beta(golden_demo):
    """Trace update with the deformation taken from the reference's run (beta after its 250 Adam steps), so that
    the only differences left are the closed-form footprint values and the accumulation of G_t, b_t: the margin
    of the statistics themselves against the reference's fp64 einsum (Demix/dNMF.py:141-148), 50 sweeps."""
    g = golden_demo
    for cutoff in (0.0, 3.5):
        dn = _demo_model(g, cutoff, resident=True)
        with torch.no_grad():
            dn.fp.beta.copy_(torch.tensor(g["beta250"]).cuda())
        dn.update_traces(None, gamma_c=0, iter_c=50)
        err = float((dn.C.cpu() - torch.tensor(g["C_mu0"])).abs().max() / np.abs(g["C_mu0"]).max())
        print("\n[traces from the reference's beta, cutoff %.1f] max |C - C_ref| / max |C_ref| = %.2e" % (cutoff, err))
        assert err <= (1e-5 if cutoff == 0.0 else 1e-4)


def test_repeated_frame_in_a_batch_sums_its_gradients(golden_random):
    """A sampler with replacement can list a frame twice in one minibatch: the reference's autograd adds both
    contributions to the frame's gradient column and the loss averages over all B entries."""
    g = golden_random
    sz = g["sz"].tolist()
    T = g["beta"].shape[2]
    e = _engine(sz, 5, T, g["pos"], g["sigma"], 0.0, (1, 1, 0, 0, 2))
    beta = torch.tensor(g["beta"]).cuda()
    C = torch.tensor(g["C"]).cuda()
    frames = torch.tensor(g["frames"])
    ids = [2, 4, 2, 1, 2]
    fb = frames[ids].clone()
    fb[2] = frames[3]                      # the second occurrence of frame 2 even carries different data
    grad, sse = e.loss_grad(torch.tensor(ids), beta, C, frames=fb.cuda())
    port = O.TorchPort(sz, 5, T, positions=g["pos"], shape_std=2.5, C0=g["C"])
    with torch.no_grad():
        port.beta.copy_(torch.tensor(g["beta"]))
    A_tC, _, _ = port.forward(ids)
    loss = torch.nn.functional.mse_loss(A_tC, fb)
    loss.backward()
    ref = port.beta.grad.numpy()
    N = int(np.prod(sz))
    assert abs(float(sse.sum()) / (len(ids) * N) - float(loss)) <= 1e-5 * float(loss)
    err = np.abs(grad.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err
    g2, s2 = e.loss_grad(torch.tensor(ids), beta, C, frames=fb.cuda())
    assert torch.equal(grad, g2) and torch.equal(sse, s2)
    # statistics are stored per frame id: a repeated id is rejected
    from dnmf_b200 import DnmfError
    with pytest.raises(DnmfError):
        e.mu_stats(torch.tensor(ids), beta, frames=fb.cuda())
    with pytest.raises(DnmfError):
        e.mu_stats(torch.tensor(ids).cuda(), beta, frames=fb.cuda())


def test_frame_ids_outside_the_slab_are_errors_not_memory_faults(golden_random):
    from dnmf_b200 import DnmfError
    g = golden_random
    sz = g["sz"].tolist()
    T = g["beta"].shape[2]
    e = _engine(sz, 5, T, g["pos"], g["sigma"], 0.0)
    beta = torch.tensor(g["beta"]).cuda()
    C = torch.tensor(g["C"]).cuda()
    frames = torch.tensor(g["frames"])[:2].cuda()
    with pytest.raises(DnmfError):                       # host ids: immediate
        e.loss_grad(torch.tensor([0, T]), beta, C, frames=frames)
    with pytest.raises(DnmfError):
        e.forward(torch.tensor([-1, 0]), beta, C)
    e.loss_grad(torch.tensor([0, T + 5]).cuda(), beta, C, frames=frames)   # device ids: asynchronous call ...
    with pytest.raises(DnmfError):
        e.check_status()                                 # ... reported at the next synchronising check
    e.check_status()                                     # cleared by reporting
    with pytest.raises(DnmfError):
        e.mu_stats(torch.tensor([0, 99]).cuda(), beta, frames=frames)
    grad, sse = e.loss_grad(torch.tensor([0, 1]), beta, C, frames=frames)  # the engine keeps working
    assert torch.isfinite(grad).all() and torch.isfinite(sse).all()
    # tensors the kernels would index blindly are validated before the call
    with pytest.raises(DnmfError):
        e.loss_grad(torch.tensor([0, 1]), beta[:, :, :3].contiguous(), C, frames=frames)
    with pytest.raises(DnmfError):
        e.loss_grad(torch.tensor([0, 1]), beta, C.cpu(), frames=frames)
    with pytest.raises(DnmfError):
        e.motion_step(torch.tensor([0, 1]), beta, torch.zeros(10, 3, T), torch.zeros_like(beta), C, 1e-3, (0.9, 0.999),
                      1e-8, 1, frames=frames)
