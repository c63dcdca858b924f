"""Build recipe for the CUDA engine (sm_100a only).

``python -m ptmcmcsampler_b200.build`` compiles the translation units under ``csrc/`` in parallel and links
them in-tree into ``ptmcmcsampler_b200/libptmcmc_b200.so`` (git-ignored; it travels to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU.  ``-DPTMCMC_ALL_SORT_CFGS`` (``--all-cfgs``) also instantiates
the experimental launch geometries of the sorted kernel (``PTMCMC_SORT_CFG``).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libptmcmc_b200.so")
SOURCES = ["engine.cu", "mh_sorted.cu", "mh_mma.cu", "user_target.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
LINK_LIBS = ["-ldl"]  # NVRTC and the driver API are dlopen()ed on first use: the library loads without a driver


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [
        os.path.join(HERE, "..", "include", "ptmcmc_b200.h")]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(deps, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def stale():
    return _newer([os.path.join(CSRC, s) for s in sources()] + headers() + [os.path.abspath(__file__)], LIB)


def build(force=False, verbose=False, all_cfgs=None):
    if all_cfgs is None:
        all_cfgs = bool(os.environ.get("PTMCMC_ALL_SORT_CFGS"))
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DPTMCMC_ALL_SORT_CFGS"] if all_cfgs else [])
    flags += os.environ.get("PTMCMC_NVCC_EXTRA", "").split()  # development aid, e.g. -DPTMCMC_MMA_CLOCKS
    hdrs = headers() + [os.path.abspath(__file__)]

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + (".all.o" if all_cfgs else ".o"))
        path = os.path.join(CSRC, src)
        if force or _newer([path] + hdrs, obj):
            out = subprocess.run([nvcc()] + flags + ["-c", "-o", obj, path], stdout=subprocess.PIPE,
                                 stderr=subprocess.STDOUT, text=True)
            if verbose or out.returncode:
                sys.stderr.write(out.stdout)
            if out.returncode:
                raise subprocess.CalledProcessError(out.returncode, out.args)
        return obj

    with ThreadPoolExecutor(max_workers=len(sources())) as ex:
        objs = list(ex.map(compile_one, sources()))
    subprocess.check_call([nvcc(), "-shared", "-o", LIB] + objs + LINK_LIBS)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, all_cfgs=("--all-cfgs" in sys.argv) or None))
