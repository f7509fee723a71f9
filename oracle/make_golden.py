"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

Run in the build container only:  python -m oracle.make_golden
The reference ships no golden vectors; these files pin the oracle and the
CUDA path to outputs of the reference's own code.  Seeds: np.random.seed(0), torch.manual_seed(0).

  demo_cfg1.npz   demo.py:16-46 configuration (K=10, T=100, 50x50x2, batch 4, Adam lr 1e-5,
                  shuffle=False): input frames, initial positions, C0, recon loss of each of the
                  first 250 Adam steps, beta after them, then C after
                  update_footprints(gamma_c=0, iter_c=50) and after a further
                  update_footprints(gamma_c=1e-2, iter_c=10); a few frames of A_t / Y_i.
  demo_shuffle.npz same data (seeds identical, so frames/pos0/C0 equal demo_cfg1.npz), DataLoader(shuffle=True)
                  with torch.manual_seed(1000+epoch) before each of 2 epochs: batch order, losses, beta.
  random_beta.npz small 3-D case with random quadratic beta (44 % out-of-bounds samples):
                  forward A_tC, A_t, loss and d loss / d beta from the reference's autograd.
  extras.npz      (python -m oracle.make_golden extras) outputs the first three files do not hold: the `reg`
                  return of ExponentialFP.forward (Demix/dNMF.py:60-61) for the random-beta case and for the
                  demo's beta after 250 steps; spatial_pushforward (Demix/dNMF.py:69-103) of the random-beta
                  case -- a 3-D volume whose deformed points are in general position, so the nearest-neighbour
                  registered video Y_i has no distance ties; the static update_spatial (Demix/dNMF.py:151-160)
                  and update_temporal (:139-149) on small dense arrays.
"""
import contextlib
import io
import os
import warnings

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

from oracle.ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def demo_cfg1(ref):
    np.random.seed(0)
    torch.manual_seed(0)
    K, T, B = 10, 100, 4
    sz = torch.tensor([50, 50, 2])
    ds = ref.SimulatedVideoDataset(K=K, T=T, sz=sz, shape_std=3, density=.2, bg_snr=-120, motion="gp",
                                   traces="exp", motion_par={"sigma": [5, 5, .01], "ls": [10, 10, 10]})
    raw_video = ds.video.clone()
    loader = DataLoader(ds, batch_size=B, shuffle=False, num_workers=0)
    dn = ref.DeformableNMF(sz, K, T, positions=ds.positions[:, :, 0])
    C0 = dn.C.clone()
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    losses = []
    for _ in range(10):                         # Demix/dNMF.py:185-191, one recon per step
        for frames, idx in loader:
            opt.zero_grad()
            with quiet():
                A_tC, _, _, _ = dn.fp(idx.tolist(), dn.C)
            rec = F.mse_loss(A_tC, frames)
            rec.backward()
            opt.step()
            losses.append(float(rec))
    beta250 = dn.fp.beta.detach().clone()
    with quiet():
        A_t, Y_i, Y = dn.update_footprints(loader, B, sz, gamma_c=0, iter_c=50)
    C_mu0 = dn.C.clone()
    with quiet():
        dn.update_footprints(loader, B, sz, gamma_c=1e-2, iter_c=10)
    C_mu1 = dn.C.clone()
    frames_all = ds.video.permute(3, 0, 1, 2).contiguous()     # clamped in place by the loader
    np.savez_compressed(
        os.path.join(OUT, "demo_cfg1.npz"),
        sz=sz.numpy(), frames=frames_all.numpy(), raw_frames0=raw_video[:, :, :, 0].numpy(),
        positions=ds.positions.numpy(), traces=np.asarray(ds.traces), pos0=ds.positions[:, :, 0].numpy(),
        C0=C0.numpy(), losses=np.asarray(losses), beta250=beta250.numpy(),
        adam_m=opt.state[dn.fp.beta]["exp_avg"].numpy(), adam_v=opt.state[dn.fp.beta]["exp_avg_sq"].numpy(),
        C_mu0=C_mu0.numpy(), C_mu1=C_mu1.numpy(),
        A_t_first4=A_t[..., :4].astype(np.float32), Y_i_first4=Y_i[..., :4].astype(np.float32),
        Y_first4=Y[..., :4].astype(np.float32))
    print("demo_cfg1: loss %.6g -> %.6g" % (losses[0], losses[-1]))


def demo_shuffle(ref):
    """demo.py:34 uses shuffle=True: two epochs with the batch order drawn from torch's global RNG,
    re-seeded right before each epoch so that the order can be replayed without the reference."""
    np.random.seed(0)
    torch.manual_seed(0)
    K, T, B = 10, 100, 4
    sz = torch.tensor([50, 50, 2])
    ds = ref.SimulatedVideoDataset(K=K, T=T, sz=sz, shape_std=3, density=.2, bg_snr=-120, motion="gp",
                                   traces="exp", motion_par={"sigma": [5, 5, .01], "ls": [10, 10, 10]})
    loader = DataLoader(ds, batch_size=B, shuffle=True, num_workers=0)
    dn = ref.DeformableNMF(sz, K, T, positions=ds.positions[:, :, 0])
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    losses, order = [], []
    for ep in range(2):
        torch.manual_seed(1000 + ep)
        for frames, idx in loader:
            opt.zero_grad()
            with quiet():
                A_tC, _, _, _ = dn.fp(idx.tolist(), dn.C)
            rec = F.mse_loss(A_tC, frames)
            rec.backward()
            opt.step()
            losses.append(float(rec))
            order.append(idx.numpy().copy())
    np.savez_compressed(os.path.join(OUT, "demo_shuffle.npz"), losses=np.asarray(losses), order=np.asarray(order),
                        beta=dn.fp.beta.detach().numpy(), C0=dn.C.numpy())
    print("demo_shuffle: loss %.6g -> %.6g" % (losses[0], losses[-1]))


def random_beta(ref):
    np.random.seed(1)
    torch.manual_seed(1)
    K, T = 5, 6
    sz = torch.tensor([20, 16, 6])
    pos = torch.rand(K, 3) * sz[None, :]
    fp = ref.ExponentialFP(sz, K, T, positions=pos, shape_std=2.5)
    with torch.no_grad():
        scale = torch.tensor([2.0, .05, .05, .05, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3])[:, None, None]
        fp.beta += scale * torch.randn(10, 3, T)
        fp.beta[:, :, 0] = torch.cat((torch.zeros(1, 3), torch.eye(3), torch.zeros(6, 3)), 0)  # one identity frame
    C = torch.rand(K, T)
    frames = torch.rand(T, *sz.tolist())
    times = list(range(T))
    with quiet():
        A_tC, A_t, grid, _ = fp(times, C)
    loss = F.mse_loss(A_tC, frames)
    loss.backward()
    oob = float(((grid.abs() > 1).any(3)).float().mean())
    np.savez_compressed(
        os.path.join(OUT, "random_beta.npz"),
        sz=sz.numpy(), pos=pos.numpy(), sigma=fp.sigma.numpy(), beta=fp.beta.detach().numpy(), C=C.numpy(),
        frames=frames.numpy(), A_tC=A_tC.detach().numpy(), A_t=A_t.detach().numpy(), loss=float(loss),
        grad=fp.beta.grad.numpy(), oob_fraction=oob)
    print("random_beta: loss %.6g, oob %.2f" % (float(loss), oob))


def extras(ref):
    g = dict(np.load(os.path.join(OUT, "random_beta.npz")))
    d = dict(np.load(os.path.join(OUT, "demo_cfg1.npz")))
    out = {}
    # -- reg of forward (Demix/dNMF.py:60-61), random quadratic beta --
    sz = torch.tensor(g["sz"])
    K, T = g["pos"].shape[0], g["beta"].shape[2]
    torch.manual_seed(5)
    fp = ref.ExponentialFP(sz, K, T, positions=torch.tensor(g["pos"]), shape_std=2.5)
    with torch.no_grad():
        fp.beta.copy_(torch.tensor(g["beta"]))
    C = torch.tensor(g["C"])
    with quiet():
        _, _, _, reg = fp(list(range(T)), C)
    out["reg_random"] = reg.detach().numpy()
    # -- reg for the demo's beta after 250 Adam steps, frames 0..7 --
    szd = torch.tensor(d["sz"])
    fpd = ref.ExponentialFP(szd, 10, 100, positions=torch.tensor(d["pos0"]))
    with torch.no_grad():
        fpd.beta.copy_(torch.tensor(d["beta250"]))
    with quiet():
        _, _, _, regd = fpd(list(range(8)), torch.tensor(d["C0"]))
    out["reg_demo250_first8"] = regd.detach().numpy()
    # -- spatial_pushforward of the random-beta case with a milder deformation (tie-free Y_i) --
    class _Model:
        pass
    m = _Model()
    m.fp, m.C = fp, C
    with torch.no_grad():
        ident = torch.cat((torch.zeros(1, 3), torch.eye(3), torch.zeros(6, 3)), 0)[:, :, None]
        fp.beta.copy_(ident + 0.25 * (torch.tensor(g["beta"]) - ident))
    out["pf_beta"] = fp.beta.detach().numpy().copy()
    frames = torch.tensor(g["frames"])
    batches = [(frames[i:i + 2], torch.arange(i, i + 2)) for i in range(0, T, 2)]
    with quiet(), torch.no_grad():
        A_t, Y_i, Y = ref.ExponentialFP.spatial_pushforward(batches, 2, sz.tolist(), "cpu", m)
    out["pf_Y_i"] = Y_i.astype(np.float32)
    out["pf_A_t_max2"] = A_t.max(2).astype(np.float32)          # demo.py:50-52 style max-projection along z
    # -- static multiplicative updates on small dense arrays --
    rs = np.random.RandomState(11)
    Kd, Td, M, N = 4, 6, 7, 5
    A = rs.rand(M, N, Kd)
    Cs = rs.rand(Kd, Td)
    Yi = rs.rand(M, N, Td)
    D = rs.rand(M, N, Kd)
    out.update(us_A=A, us_C=Cs, us_Yi=Yi, us_D=D, us_gamma=np.float64(0.7),
               us_out_D=ref.DeformableNMF.update_spatial(A.copy(), Cs, Yi, D=D, gamma=0.7),
               us_out_noD=ref.DeformableNMF.update_spatial(A.copy(), Cs, Yi))
    At = rs.rand(5, 4, 3, Kd, Td)
    Yt = rs.rand(5, 4, 3, Td)
    out.update(ut_A_t=At, ut_Y=Yt, ut_out_none=ref.DeformableNMF.update_temporal(At, Cs.copy(), Yt),
               ut_out_gamma=ref.DeformableNMF.update_temporal(At, Cs.copy(), Yt, gamma=1e-2))
    np.savez_compressed(os.path.join(OUT, "extras.npz"), **out)
    print("extras: reg_random", out["reg_random"], "reg_demo", out["reg_demo250_first8"][:3])


def main():
    import sys
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    ref, _ = load_reference()
    if "extras" in sys.argv[1:]:
        extras(ref)
        return
    demo_cfg1(ref)
    demo_shuffle(ref)
    random_beta(ref)
    extras(ref)


if __name__ == "__main__":
    main()
