"""Does the fused pass give the same bits for a frame wherever it sits in a CTA's chunk of frames?  Gradient and SSE of
125 cfg2 frames (identity and a deformation per frame) with DNMF_FPC = 1, 3, 8 and with / without the short-tail split."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import bench
    cfg = bench.CONFIGS["cfg2"]
    T = 125
    dev = torch.device("cuda:0")
    dn, vid = bench.build_model(cfg, T, dev, 1, None)
    eng = dn.fp.engine
    beta = dn.fp.beta.detach()
    ids = torch.arange(T, dtype=torch.int32, device=dev)
    out = {}
    for state in ("identity", "deformed"):
        if state == "deformed":
            bench.deform(beta, dn.affine, T, dev)
        g, sse = eng.loss_grad(ids, beta, dn.C)
        yh = eng.forward(ids[:24], beta, dn.C)[0] if state == "identity" else None
        out[state] = (g.cpu(), sse.cpu(), None if yh is None else yh.cpu())
    torch.save(out, sys.argv[2])
    sys.exit(0)

ref = None
for label, env in (("fpc1", {"DNMF_FPC": "1"}), ("fpc1 again", {"DNMF_FPC": "1"}), ("fpc3", {"DNMF_FPC": "3", "DNMF_FPC_TAIL_OFF": "1"}),
                   ("fpc8", {"DNMF_FPC": "8", "DNMF_FPC_TAIL_OFF": "1"})):
    e = dict(os.environ)
    e.update(env)
    path = "/tmp/chunk_bits_%s.pt" % label
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", path], env=e, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cur = torch.load(path)
    if ref is None:
        ref = cur
        print(label, "reference")
        continue
    for state in ("identity", "deformed"):
        g0, s0, y0 = ref[state]
        g1, s1, y1 = cur[state]
        print("   frames whose sse differs:", torch.nonzero(s0 != s1).flatten().tolist())
        if y0 is not None:
            d = torch.nonzero(y0 != y1)
            print("   Yhat voxels that differ: %d; first: %s" % (d.shape[0], d[:12].tolist()))
        nd = int((g0 != g1).sum())
        print("%s %s: gradient entries that differ %d of %d (max rel %.2e), sse entries that differ %d"
              % (label, state, nd, g0.numel(), float(((g0 - g1).abs() / g0.abs().clamp_min(1e-30)).max()), int((s0 != s1).sum())))
