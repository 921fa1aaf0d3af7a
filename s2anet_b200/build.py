"""Compile the sm_100a CUDA sources into s2anet_b200/csrc/libs2a_b200.so (in-tree, no JIT cache).

    python -m s2anet_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the gpurun
snapshot.  Objects are rebuilt only when a source or header is newer.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libs2a_b200.so")
OBJ = os.path.join(CSRC, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, verbose):
    extra = os.environ.get("S2A_NVCC_EXTRA", "").split()      # e.g. -DS2A_TC_TIMELINE (in-kernel clock probes)
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _newer(o, [s] + hdrs):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(lambda j: _compile(j[0], j[1], verbose), jobs):
                if verbose and out:
                    print(out)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if jobs or _newer(LIB, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-soname=libs2a_b200.so", "-o", LIB] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s" % r.stdout)
    return LIB


TORCH_EXT = os.path.join(HERE, "_s2a_torch.so")
TORCH_EXT_SRC = os.path.join(CSRC, "torch_binding.cpp")


def build_torch_ext(force=False):
    """g++ the thin torch-extension binding (csrc/torch_binding.cpp: at::Tensor in, one C-ABI call, at::Tensor out) into
    s2anet_b200/_s2a_torch.so, linked against csrc/libs2a_b200.so (rpath $ORIGIN/csrc) and torch's own libraries.  In-tree
    like the CUDA library, so it travels with the gpurun snapshot.  No CUDA sources: builds without a GPU."""
    hdrs = glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not (force or _newer(TORCH_EXT, [TORCH_EXT_SRC, LIB] + hdrs)):
        return TORCH_EXT
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    try:                                                   # (the argument is a device-type string in recent torch, a flag before)
        incs, libdirs = ce.include_paths("cuda"), ce.library_paths("cuda")
    except TypeError:
        incs, libdirs = ce.include_paths(True), ce.library_paths(True)
    incs = [os.path.join(ROOT, "include")] + incs + [os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"]]
    libdirs = libdirs + [os.path.join(cuda_home, "lib64")]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_s2a_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-Wno-deprecated-declarations"]
    for i in incs:
        cmd += ["-isystem" if "site-packages" in i or "cuda" in i else "-I", i]
    cmd += [TORCH_EXT_SRC, "-o", TORCH_EXT, "-L", CSRC, "-l:libs2a_b200.so", "-Wl,-rpath,$ORIGIN/csrc"]
    for d in libdirs:
        cmd += ["-L", d, "-Wl,-rpath," + d]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for %s:\n%s" % (TORCH_EXT_SRC, r.stdout))
    return TORCH_EXT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
