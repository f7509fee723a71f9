"""GPU tests of the callers and data formats either side of the fit path: the static
multiplicative updates on dense arrays, the synthetic generator kernel, the lazy pushforward and the z
max-projections.  All through the C ABI; oracle = numpy restatements pinned to the reference's own outputs
(tests/golden/extras.npz, tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import dnmf_oracle as O

pytestmark = pytest.mark.gpu


def test_update_spatial_matches_reference(golden_extras):
    """DeformableNMF.update_spatial (Demix/dNMF.py:151-160) against the real reference's output, with and without
    the distance penalty D."""
    from dnmf_b200 import DeformableNMF
    x = golden_extras
    got = DeformableNMF.update_spatial(x["us_A"], x["us_C"], x["us_Yi"], D=x["us_D"], gamma=float(x["us_gamma"]))
    np.testing.assert_allclose(got, x["us_out_D"], rtol=1e-12)
    got = DeformableNMF.update_spatial(x["us_A"], x["us_C"], x["us_Yi"])
    np.testing.assert_allclose(got, x["us_out_noD"], rtol=1e-12)


@pytest.mark.parametrize("shape,K,T", [((9, 7), 5, 11), ((40, 30, 6), 37, 70), ((130, 3), 33, 64)])
def test_update_spatial_ragged_sizes_and_on_the_fly_penalty(shape, K, T):
    from dnmf_b200.engine import update_spatial_dense
    rs = np.random.RandomState(sum(shape) + K)
    A = rs.rand(*shape, K)
    C = rs.rand(K, T)
    Yi = rs.rand(*shape, T)
    D = rs.rand(*shape, K)
    ref = O.update_spatial(A, C, Yi, D=D, gamma=0.3)
    got = update_spatial_dense(A, C, Yi, D=D, gamma=0.3).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    if len(shape) == 3:   # D computed in the kernel from the positions (Demix/dNMF.py:133-135), never stored
        pos = (rs.rand(K, 3) * np.asarray(shape)).astype(np.float32)
        ref = O.update_spatial(A, C, Yi, D=O.distance_penalty(shape, pos), gamma=1.0)
        got = update_spatial_dense(A, C, Yi, gamma=1.0, positions=pos, grid=shape).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-11)


def test_update_temporal_static_matches_reference(golden_extras):
    from dnmf_b200 import DeformableNMF
    x = golden_extras
    got = DeformableNMF.update_temporal(x["ut_A_t"], x["us_C"], x["ut_Y"])
    np.testing.assert_allclose(got, x["ut_out_none"], rtol=1e-12)
    got = DeformableNMF.update_temporal(x["ut_A_t"], x["us_C"], x["ut_Y"], gamma=1e-2)
    np.testing.assert_allclose(got, x["ut_out_gamma"], rtol=1e-12)
    # repeated calls are bitwise identical (fixed summation order)
    again = DeformableNMF.update_temporal(x["ut_A_t"], x["us_C"], x["ut_Y"], gamma=1e-2)
    assert np.array_equal(got, again)


@pytest.mark.parametrize("sz,K,T,shape_std", [([50, 50, 2], 10, 5, 3.0), ([37, 21, 9], 7, 3, 3.0),
                                                ([64, 48, 21], 150, 2, 18.0), ([16, 16, 64], 3, 2, 5.0)])
def test_generator_kernel_matches_the_separable_cells(sz, K, T, shape_std):
    """dnmf_render_cells against the fp64 restatement of the Simulator's cells (WUtils/Simulator.py:66-73,197-203),
    ragged volumes, more neurons than one staged chunk, neurons far outside the volume."""
    from dnmf_b200.engine import render_cells
    rs = np.random.RandomState(K + T)
    pos = (rs.rand(K, 3, T) * np.asarray(sz)[None, :, None]).astype(np.float32)
    pos[0] = np.asarray([-80.0, 5.0, 1.0])[:, None]            # never reaches the volume
    tr = (1 + rs.rand(K, T)).astype(np.float32)
    got = render_cells(torch.tensor(pos), tr, sz, shape_std).cpu().numpy()
    ref = O.render_cells(pos, tr, sz, shape_std)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6 * ref.max())


def test_generator_kernel_is_bitwise_reproducible():
    """Every neuron of a staged chunk has its own slot and the columns add the live ones in ascending k: rendering the
    same cells twice (many neurons per tile, several chunks) gives the same bits."""
    from dnmf_b200.engine import render_cells
    sz, K, T = [48, 40, 9], 230, 3
    rs = np.random.RandomState(5)
    pos = (rs.rand(K, 3, T) * np.asarray(sz)[None, :, None]).astype(np.float32)
    tr = (1 + rs.rand(K, T)).astype(np.float32)
    first = render_cells(torch.tensor(pos), tr, sz, 9.0)
    for _ in range(5):
        assert torch.equal(first, render_cells(torch.tensor(pos), tr, sz, 9.0))


def test_generate_video_on_the_gpu_uses_the_kernel_and_matches_the_cpu_path():
    from dnmf_b200.simulate import generate_video, render_clean
    args = (6, 4, [24, 20, 5], 3, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]})
    vg, pg, tg = generate_video(*args, seed=3, device="cuda", frame_major=True)
    vc, pc, tc = generate_video(*args, seed=3, device="cpu", frame_major=True)
    assert torch.equal(pg, pc) and np.array_equal(tg, tc)
    clean_g = render_clean(pg, tg, [24, 20, 5], 3.0, device="cuda")
    clean_c = render_clean(pc, tc, [24, 20, 5], 3.0, device="cpu")
    np.testing.assert_allclose(clean_g.cpu().numpy(), clean_c.numpy(), rtol=2e-5, atol=1e-6)
    assert vg.shape == vc.shape and float(vg.max()) == 1.0


def test_lazy_pushforward_and_max_projections(golden_random, golden_extras):
    """pushforward_chunks / max_projections against the reference's dense spatial_pushforward outputs
    (tests/golden/extras.npz): A_t.max(2), Y_i.max(2), Y.max(2) as demo.py:50-52 takes them."""
    from dnmf_b200 import DeformableNMF, ExponentialFP, FrameDataset
    g, x = golden_random, golden_extras
    sz = g["sz"].tolist()
    T = x["pf_beta"].shape[2]
    dn = DeformableNMF(sz, 5, T, positions=torch.tensor(g["pos"]), shape_std=2.5, cutoff=0.0, verbose=False)
    with torch.no_grad():
        dn.fp.beta.copy_(torch.tensor(x["pf_beta"]).cuda())
    dn.C = torch.tensor(g["C"]).cuda()
    frames = torch.tensor(g["frames"])
    loader = torch.utils.data.DataLoader(FrameDataset(frames), batch_size=2, shuffle=False)
    A_max, Yi_max, Y_max = ExponentialFP.max_projections(loader, dn)
    np.testing.assert_allclose(A_max.cpu().numpy(), x["pf_A_t_max2"], atol=2e-6)
    assert np.array_equal(Y_max.cpu().numpy(), frames.numpy().max(3).transpose(1, 2, 0))
    ref_yi = x["pf_Y_i"].max(2)
    same = float((Yi_max.cpu().numpy() == ref_yi).mean())
    assert same > 0.99, same           # the handful of distance ties may pick the other source (see test_gpu_parity)
    n = 0
    for ids, At, yi, y in ExponentialFP.pushforward_chunks(loader, dn):
        assert At.shape == (2, 5, *sz) and yi.shape == (2, *sz) and torch.equal(y.cpu(), frames[ids.long()])
        np.testing.assert_allclose(At.amax(4).cpu().numpy(), np.moveaxis(x["pf_A_t_max2"][..., ids.long().numpy()], (2, 3), (1, 0)),
                                   atol=2e-6)
        n += 1
    assert n == T // 2
