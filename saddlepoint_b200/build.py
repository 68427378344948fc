"""In-tree build of libsaddlepoint_b200.so (nvcc, sm_100a only).

Every translation unit is compiled to its own object (in parallel, only when stale) and the
objects are linked into saddlepoint_b200/lib/libsaddlepoint_b200.so."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libsaddlepoint_b200.so")
SOURCES = ["context.cu", "assembly.cu", "pcg.cu", "pcg_peer.cu", "lm.cu", "batch.cu", "problems.cu", "cholesky.cu", "dist.cu",
           "symbolic.cpp", "pairs_v2.cpp", "assembly_v2.cu", "symbolic_api.cpp", "analysis_update.cu", "chol_symbolic.cpp", "ic0_symbolic.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".hpp", ".cuh"))]
    hs.append(os.path.join(_HERE, "..", "include", "saddlepoint_b200.h"))
    return hs


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into saddlepoint_b200/lib/ (no-op if fresh)."""
    if not force and not _stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdr_t = max(os.path.getmtime(hp) for hp in _headers())
    jobs = []
    objs = []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for rc, cmd in zip(ex.map(lambda c: subprocess.run(c).returncode, jobs), jobs):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, cmd)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
