"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no
compute without a GPU), frame sharding / halo exchange under gloo with world_size 2, simulator."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import dnmf_b200
    names = dnmf_b200.declared_symbols()
    assert len(names) >= 25 and "dnmf_motion_step" in names and "dnmf_bin_tiles" in names
    lib = ctypes.CDLL(dnmf_b200.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    lib2 = dnmf_b200.load()
    assert lib2.dnmf_abi_version() == 7
    for n in names:                                   # every declared symbol has a ctypes signature
        assert n in lib2._signatures, n


def test_checked_build_exists_and_says_so():
    """`python -m dnmf_b200.build --checked` (built by __graft_entry__.build()) is the same library with device-side
    assertions compiled in; dnmf_build_info tells the two apart, and both export every declared symbol."""
    import dnmf_b200
    from dnmf_b200 import build as b
    checked = ctypes.CDLL(b.build(checked=True))
    normal = ctypes.CDLL(b.build())
    for lib in (checked, normal):
        lib.dnmf_build_info.restype = ctypes.c_int
        for n in dnmf_b200.declared_symbols():
            assert hasattr(lib, n), n
    assert checked.dnmf_build_info() & 1 == 1
    assert normal.dnmf_build_info() & 1 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import dnmf_b200
    lib = dnmf_b200.load()
    h = ctypes.c_void_p()
    rc = lib.dnmf_create(ctypes.byref(h), 8, 8, 2, 3, 4, 0)
    assert rc != 0 and b"no CPU fallback" in lib.dnmf_last_error()
    with pytest.raises(dnmf_b200.DnmfError):
        dnmf_b200.Engine([8, 8, 2], 3, 4)
    with pytest.raises(dnmf_b200.DnmfError):
        dnmf_b200.DeformableNMF([8, 8, 2], 3, 4)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dnmf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_frame_slab_partition():
    from dnmf_b200.sharding import frame_slab, owner_of, split_batch
    for T, W in ((100, 8), (5000, 8), (7, 3), (4, 8), (1000, 1)):
        seen = []
        for r in range(W):
            s, c = frame_slab(T, W, r)
            seen += list(range(s, s + c))
            for t in range(s, s + c):
                assert owner_of(t, T, W) == r
        assert seen == list(range(T))
        sizes = [frame_slab(T, W, r)[1] for r in range(W)]
        assert max(sizes) - min(sizes) <= 1
    mine, B = split_batch([3, 50, 51, 99], 100, 2, 1)
    assert mine == [0, 1, 49] and B == 4


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dnmf_b200.sharding import frame_slab, make_halo_exchange, allreduce_loss
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
K = 5
start, count = frame_slab(11, world, rank)
C = torch.arange(11 * K, dtype=torch.float64).reshape(11, K)        # global traces [T][K]
mine = C[start:start + count]
prev, nxt = make_halo_exchange()(mine[0].clone(), mine[-1].clone())
if rank == 0:
    assert prev is None and torch.equal(nxt, C[start + count])
else:
    assert nxt is None and torch.equal(prev, C[start - 1])
loss = allreduce_loss(torch.tensor([float(rank + 1)], dtype=torch.float64))
assert float(loss) == 3.0
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_halo_exchange_and_loss_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_simulator_cells_match_reference_frames(golden_demo):
    """render_clean reproduces the reference Simulator's noise-free cells: after the Simulator's own
    normalisation the golden (noisy, bg_snr=-120 dB) frame differs from ours by the noise only."""
    from dnmf_b200.simulate import render_clean
    g = golden_demo
    sz = g["sz"].tolist()
    clean = render_clean(torch.tensor(g["positions"]), g["traces"], sz, 3.0, device="cpu")   # [T,X,Y,Z]
    raw0 = torch.tensor(g["raw_frames0"])                                                     # reference, t=0
    c0 = clean[0]
    # the reference frame is a*clean + noise with one global scale a: fit it and check the residual is white
    a = float((raw0 * c0).sum() / (c0 * c0).sum())
    resid = raw0 - a * c0
    assert resid.std() < 0.05 * float(raw0.max())
    assert abs(float((resid * c0).sum())) < 1e-3 * float((c0 * c0).sum()) * abs(a)
    corr = float(torch.corrcoef(torch.stack((raw0.flatten(), c0.flatten())))[0, 1])
    assert corr > 0.97


def test_generate_video_shapes_and_determinism():
    from dnmf_b200.simulate import SimulatedVideoDataset, generate_video
    v1, p1, t1 = generate_video(4, 6, [16, 12, 3], 3, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]},
                                seed=7, device="cpu")
    v2, p2, t2 = generate_video(4, 6, [16, 12, 3], 3, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]},
                                seed=7, device="cpu")
    assert v1.shape == (16, 12, 3, 6) and p1.shape == (4, 3, 6) and t1.shape == (4, 6)
    assert torch.equal(v1, v2) and torch.equal(p1, p2) and np.array_equal(t1, t2)
    assert float(v1.max()) == 1.0 and t1.min() >= 1.0
    ds = SimulatedVideoDataset(4, 6, [16, 12, 3], 3, .2, -120, "exp", "gp", {"sigma": [5, 5, .01], "ls": [10, 10, 10]}, seed=7)
    frame, idx = ds[2]
    assert frame.shape == (16, 12, 3) and idx == 2 and float(frame.min()) >= 0 and len(ds) == 6
    assert ds.video.shape == (16, 12, 3, 6)


def test_bench_reference_arm_runs_small():
    """--impl reference prints one JSON line with the contract's keys (tiny config so CI stays fast)."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_neuropal_dataset_from_mat(tmp_path):
    """Demix/dNMF.py:220-248 semantics on a synthetic data.mat / traces_n.mat pair."""
    from scipy.io import savemat
    from dnmf_b200 import NeuroPALVideoDataset
    rng = np.random.default_rng(0)
    data = rng.normal(size=(12, 10, 20, 7)).astype(np.float32)
    positions = 1 + rng.random((3, 3, 7)) * np.array([12, 10, 20])[None, :, None]
    savemat(tmp_path / "data.mat", {"data": data})
    savemat(tmp_path / "traces_n.mat", {"positions": positions, "neuron_names": np.array(["AVA", "AVB", "RIM"], dtype=object)})
    ds = NeuroPALVideoDataset(str(tmp_path), frames=5)
    assert len(ds) == 5 and ds.video.shape == (6, 5, 2, 5)
    frame, idx = ds[3]
    np.testing.assert_array_equal(frame.numpy(), np.clip(data[::2, ::2, ::10, 3], 0, None))
    np.testing.assert_allclose(ds.positions[:, 0, :].numpy(), (positions[:, 0, :] - 1) / 2, rtol=1e-6)
    np.testing.assert_allclose(ds.positions[:, 2, :].numpy(), (positions[:, 2, :] - 1) / 10, rtol=1e-6)


def test_loader_index_batches_match_dataloader_iteration():
    """Walking a DataLoader's sampler for the minibatch ids (attached video: no frame is loaded) yields the same
    ids in the same order as iterating the DataLoader, consumes the global random stream identically (shuffled
    epochs stay reproducible against the reference under one torch.manual_seed), and never touches the dataset."""
    import numpy as np
    import torch
    from torch.utils.data import DataLoader, Dataset
    from dnmf_b200.model import loader_index_batches
    from dnmf_b200.simulate import FrameDataset

    class Untouchable(Dataset):
        returns_frame_index = True

        def __len__(self):
            return 23

        def __getitem__(self, idx):
            raise AssertionError("the dataset must not be read")

    real = FrameDataset(torch.zeros(23, 2, 2, 1))
    for kw in (dict(batch_size=4, shuffle=True), dict(batch_size=4, shuffle=False),
               dict(batch_size=5, shuffle=True, drop_last=True)):
        torch.manual_seed(123)
        ref = [[np.asarray(d[1]).reshape(-1).tolist() for d in DataLoader(real, **kw)] for _ in range(3)]
        state_ref = torch.get_rng_state()
        torch.manual_seed(123)
        loader = DataLoader(Untouchable(), **kw)
        got = [[b.tolist() for b in loader_index_batches(loader)] for _ in range(3)]
        assert got == ref
        assert torch.equal(torch.get_rng_state(), state_ref)
    g = torch.Generator().manual_seed(5)
    ref = [np.asarray(d[1]).tolist() for d in DataLoader(real, batch_size=3, shuffle=True, generator=g)]
    g2 = torch.Generator().manual_seed(5)
    got = [b.tolist() for b in loader_index_batches(DataLoader(Untouchable(), batch_size=3, shuffle=True, generator=g2))]
    assert got == ref and torch.equal(g.get_state(), g2.get_state())
    # not a DataLoader, or a dataset that does not declare (frame, index) items: no shortcut
    assert loader_index_batches([(None, torch.arange(4))]) is None
    assert loader_index_batches(DataLoader(torch.utils.data.TensorDataset(torch.zeros(8, 1)), batch_size=2)) is None


def test_collect_id_batches_from_dataloaders_and_plain_iterables():
    """ids + batch offsets handed to dnmf_motion_epoch / dnmf_mu_stats: the same from a DataLoader (sampler walk),
    from a DataLoader over an unmarked dataset (items are read) and from a plain list of (frames, ids) items;
    ragged last batch; empty loader."""
    import numpy as np
    import torch
    from torch.utils.data import DataLoader, Dataset
    from dnmf_b200.model import collect_id_batches
    from dnmf_b200.simulate import FrameDataset

    class Unmarked(Dataset):
        def __len__(self):
            return 10

        def __getitem__(self, idx):
            return torch.zeros(1), idx

    for ds in (FrameDataset(torch.zeros(10, 1, 1, 1)), Unmarked()):
        ids, off = collect_id_batches(DataLoader(ds, batch_size=4, shuffle=False))
        assert ids.dtype == np.int32 and off.dtype == np.int32 and ids.flags["C_CONTIGUOUS"]
        assert ids.tolist() == list(range(10)) and off.tolist() == [0, 4, 8, 10]
    items = [(None, torch.tensor([3, 1])), (None, np.array([7])), (None, [2, 0, 5])]
    ids, off = collect_id_batches(items)
    assert ids.tolist() == [3, 1, 7, 2, 0, 5] and off.tolist() == [0, 2, 3, 6]
    ids, off = collect_id_batches([])
    assert ids.size == 0 and off.tolist() == [0]


def test_update_motion_and_traces_host_plumbing_with_a_recording_engine(monkeypatch):
    """Host logic of the attached-video path without a GPU: a recording stand-in for the ctypes engine checks
    what `update_motion` / `update_footprints` hand to dnmf_motion_epoch and dnmf_mu_stats when the loaders are
    torch DataLoaders (sampler walk: the dataset is never read) -- ids, batch offsets, Adam step numbering over
    epochs, loss bookkeeping."""
    import numpy as np
    import torch
    from torch.utils.data import DataLoader, Dataset
    import dnmf_b200.model as M

    calls = []

    class RecordingEngine:
        def __init__(self, sz, K, T, device=None):
            self.device = torch.device("cpu")
            self.X, self.Y, self.Z = sz
            self.K, self.T, self.N = K, T, int(np.prod(sz))

        def set_tiling(self, *a):
            pass

        def set_affine(self, affine):
            calls.append(("set_affine", bool(affine)))

        def set_footprints(self, *a):
            pass

        def upload_frames(self, frames, t0=0, clamp_negative=True):
            calls.append(("upload", tuple(frames.shape)))

        def motion_epoch(self, ids_dev, offsets, beta, m, v, C, lr, betas, eps, first_step, affine=False,
                         global_batch_scale=1, loss_out=None):
            assert ids_dev.dtype == torch.int32 and np.asarray(offsets).dtype == np.int32
            calls.append(("epoch", ids_dev.tolist(), np.asarray(offsets).tolist(), int(first_step), float(lr)))
            loss_out.copy_(torch.arange(len(offsets) - 1, dtype=torch.float64) + first_step)

        def mu_stats(self, ids, beta, frames=None):
            assert frames is None
            calls.append(("stats", ids.tolist()))

        def mu_sweeps(self, C, gamma, iters):
            calls.append(("sweeps", gamma, iters))

        def check_status(self):
            calls.append(("check",))

    class NeverRead(Dataset):
        returns_frame_index = True

        def __len__(self):
            return 10

        def __getitem__(self, idx):
            raise AssertionError("an attached video makes the loader's frames unnecessary")

    monkeypatch.setattr(M, "Engine", RecordingEngine)
    sz, K, T = [6, 5, 2], 3, 10
    dn = M.DeformableNMF(sz, K, T, positions=np.ones((K, 3), np.float32), verbose=False)
    dn.attach_video(torch.zeros(T, *sz), layout="TXYZ")
    opt = torch.optim.Adam([dn.fp.beta], lr=1e-5)
    loader = DataLoader(NeverRead(), batch_size=4, shuffle=False)
    dn.update_motion(loader, opt, gamma=1, epochs=2)
    epochs = [c for c in calls if c[0] == "epoch"]
    assert len(epochs) == 2
    assert epochs[0][1] == list(range(10)) and epochs[0][2] == [0, 4, 8, 10]
    assert epochs[0][3] == 1 and epochs[1][3] == 4                  # 3 minibatches per epoch: Adam steps 1-3, 4-6
    assert abs(epochs[0][4] - 1e-5) < 1e-12
    assert int(opt.state[dn.fp.beta]["step"]) == 6
    assert dn.losses().tolist() == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    dn.update_footprints(loader, 4, sz, gamma_c=0, iter_c=50, dense=False)
    stats = [c for c in calls if c[0] == "stats"]
    assert sum((c[1] for c in stats), []) == list(range(10))
    assert calls[-1] == ("sweeps", 0, 50)
    assert ("check",) in calls                                      # update_motion ends with the device status check
