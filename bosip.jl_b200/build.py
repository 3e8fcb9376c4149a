"""Build libbosip_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python bosip.jl_b200/build.py [--force]

Each .cu is compiled to an object in csrc/build/ (in parallel, skipped when up to date) and linked into
bosip.jl_b200/libbosip_b200.so, which travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libbosip_b200.so")
# (source, object stem, extra flags): kstar.cu is compiled once per kernel family, its 165 template instantiations dominate the build
SOURCES = [("kstar.cu", f"kstar_k{k}", [f"-DKSTAR_KERNEL_ID={k}"]) for k in range(3)] + [
    (f, f[:-3], []) for f in ("api.cu", "vargemm.cu", "vargemm_i8.cu", "epilogue.cu", "fit.cu", "metrics.cu", "sampler.cu", "peaks.cu", "multi.cu", "confset.cu")]
HEADERS = ["common.cuh", "philox.cuh", "fastmath.cuh", os.path.join("..", "..", "include", "bosip_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]

    def compile_one(item):
        src, stem, extra = item
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, stem + ".o")
        if not force and not _stale(o, [s] + hdrs):
            return o, ""
        r = subprocess.run([NVCC, *FLAGS, *extra, "-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {extra}:\n{r.stdout}\n{r.stderr}")
        with open(o + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        return o, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
