"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m dnmf_b200.build            # -> dnmf_b200/_C/libdnmf_b200.so
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "dnmf_kernels.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "dnmf_device.cuh"), os.path.join(HERE, "csrc", "dnmf_mu.inc.cu"), os.path.join(HERE, "csrc", "dnmf_ext.inc.cu"),
        os.path.join(os.path.dirname(HERE), "include", "dnmf_b200.h")]
OUT_DIR = os.path.join(HERE, "_C")
OUT = os.path.join(OUT_DIR, "libdnmf_b200.so")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--fmad=true", "-Xcompiler", "-fPIC,-O2", "-shared", "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
