"""Import shim for the REAL reference (test infrastructure only, never product code).

The reference at /root/reference cannot be imported as shipped: Demix/dNMF.py:7 imports
`Methods.Demix.WUtils`, Demix/dNMF.py:16 hard-codes device='cuda', and demo.py needs
matplotlib.  This shim registers the missing package names, points `device` at the CPU and
returns the module.  It only works where /root/reference exists (the build container); the
GPU box never has it, so nothing under tests -m gpu / bench.py / smoke() may call this.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DNMF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Demix", "dNMF.py"))


def load_reference():
    """Returns (ref_module, simulator_module) with ref_module.device == 'cpu'."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import WUtils.Simulator as sim  # noqa: E402
    for name in ("Methods", "Methods.Demix", "Methods.Demix.WUtils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["Methods.Demix.WUtils"].Simulator = sim
    import Demix.dNMF as ref  # noqa: E402
    ref.device = "cpu"
    return ref, sim
