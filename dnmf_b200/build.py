"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m dnmf_b200.build [--force] [-v]      # -> dnmf_b200/_C/libdnmf_b200.so
    python -m dnmf_b200.build --checked           # -> dnmf_b200/_C/libdnmf_b200_checked.so  (-DDNMF_CHECKED: device-side
                                                  #    assertions on every unclamped index; tests/test_gpu_checked.py)

Every translation unit under csrc/ is compiled to its own object (in parallel; an object is rebuilt only when
its source or one of the headers is newer) and the objects are linked into one shared library.  The fused
kernel's template instantiations are spread over fit_mode<N>.cu for that reason.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = ["dnmf_kernels.cu", "fit_mode0.cu", "fit_mode1.cu", "fit_mode2.cu", "fit_mode3.cu", "dnmf_gram_tc.cu",
         "dnmf_aux.cu"]
OUT_DIR = os.environ.get("DNMF_B200_OUT_DIR") or os.path.join(HERE, "_C")   # kernel experiments build elsewhere
OBJ_DIR = os.path.join(OUT_DIR, "obj")
OUT = os.path.join(OUT_DIR, "libdnmf_b200.so")
OUT_CHECKED = os.path.join(OUT_DIR, "libdnmf_b200_checked.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "--fmad=true", "-Xcompiler", "-fPIC,-O2"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc.cu"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "dnmf_b200.h"))
    return hs


def _units():
    return [u for u in UNITS if os.path.isfile(os.path.join(CSRC, u))]


def _obj(unit: str, checked: bool = False) -> str:
    return os.path.join(OBJ_DIR + ("_checked" if checked else ""), os.path.splitext(unit)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def up_to_date(checked: bool = False) -> bool:
    hs = _headers()
    srcs = [os.path.join(CSRC, u) for u in _units()]
    return not _stale(OUT_CHECKED if checked else OUT, srcs + hs)


def _compile(unit: str, verbose: bool, checked: bool = False) -> str:
    cmd = [nvcc_path()] + ARCH + FLAGS + (["-DDNMF_CHECKED"] if checked else []) + \
          ["-c", "-o", _obj(unit, checked), os.path.join(CSRC, unit)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    extra = os.environ.get("DNMF_NVCC_FLAGS")
    if extra:
        cmd[1:1] = extra.split()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (unit, res.stdout, res.stderr))
    return res.stderr


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    out = OUT_CHECKED if checked else OUT
    if not force and up_to_date(checked):
        return out
    os.makedirs(os.path.dirname(_obj("x.cu", checked)), exist_ok=True)
    hs = _headers()
    todo = [u for u in _units() if force or _stale(_obj(u, checked), [os.path.join(CSRC, u)] + hs)]
    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as pool:
        for u, log in zip(todo, pool.map(lambda u: _compile(u, verbose, checked), todo)):
            if verbose:
                sys.stderr.write("== %s\n%s" % (u, log))
    cmd = [nvcc_path()] + ARCH + ["-shared", "-o", out] + [_obj(u, checked) for u in _units()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
