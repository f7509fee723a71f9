"""dnmf_b200 -- B200-native implementation of the dNMF fit hot path (reference: mathdiane/dNMF).

The CUDA library (dnmf_b200/_C/libdnmf_b200.so, C ABI in include/dnmf_b200.h) does all the
arithmetic; this package mirrors the reference's Python class surface on top of it.
"""
from ._lib import DnmfError, LIB_PATH, declared_symbols, load  # noqa: F401
from .simulate import FrameDataset, NeuroPALVideoDataset, SimulatedVideoDataset, generate_video  # noqa: F401


def __getattr__(name):
    # model/engine import torch.cuda-facing code lazily so CPU-only tooling can import the package
    if name in ("ExponentialFP", "DeformableNMF"):
        from . import model
        return getattr(model, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)

__version__ = "0.1.0"
