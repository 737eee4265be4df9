"""Build libspeinet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m speinet_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libspeinet_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [os.path.join(PKG, "..", "include", "speinet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_link(out: str, defines, objdir: str):
    """Compile every .cu to an object in parallel (one nvcc each), then link the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run([nvcc] + cflags + ["-c", src, "-o", obj], capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(one, sources()))
    log = "".join(r.stdout + r.stderr for _, r in results)
    if any(r.returncode != 0 for _, r in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed while compiling " + ", ".join(os.path.basename(o) for o, r in results if r.returncode != 0))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [o for o, _ in results],
                         capture_output=True, text=True)
    log += res.stdout + res.stderr
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc link failed")
    return log


def build_variant(out: str, defines, verbose: bool = False) -> str:
    """An A/B build of the same sources with extra -D defines (e.g. SPEI_TOPK=8) into `out`;
    select it at run time with SPEINET_B200_LIB=<out>."""
    log = _compile_link(out, defines, os.path.join(PKG, "build", "obj_" + os.path.basename(out)))
    if verbose:
        sys.stderr.write(log)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    log = _compile_link(LIB, [], os.path.join(PKG, "build", "obj"))
    if verbose:
        sys.stderr.write(log)
    with open(os.path.join(PKG, "build_ptxas.log"), "w") as f:
        f.write(log)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m speinet_b200.build --variant out.so SPEI_TOPK=8 ...
        i = sys.argv.index("--variant")
        print(build_variant(os.path.abspath(sys.argv[i + 1]), sys.argv[i + 2:], verbose=True))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
