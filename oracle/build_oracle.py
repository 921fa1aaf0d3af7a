"""Build the CPU oracle (TEST INFRASTRUCTURE) and, when /root/reference is present, oracle/_ref.

    python oracle/build_oracle.py            # libs2a_oracle.so (+ _ref if the reference is mounted)

* oracle/libs2a_oracle.so   <- oracle/s2a_oracle.c        (gcc, -ffp-contract=off)
* oracle/_ref/libref_iou_*.so <- oracle/ref_shim_iou.cpp + the reference header, compiled where
  it lies under /root/reference (g++ = CPU semantics, nvcc host pass = CUDA semantics).
* oracle/_ref/<ext>/<ext>.so  <- the reference's own torch extensions (CPU builds, and CUDA
  builds for sm_100a) compiled from the sources in /root/reference by torch's cpp_extension;
  see build_ref_extensions().  Nothing from the reference is copied into the repo.

The built files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
REF_OUT = os.path.join(HERE, "_ref")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def build_oracle(force=False):
    srcs = [os.path.join(HERE, "s2a_oracle.c"), os.path.join(HERE, "poly_oracle.c")]
    out = os.path.join(HERE, "libs2a_oracle.so")
    if force or _stale(out, srcs):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
              "-fvisibility=hidden", "-o", out] + srcs + ["-lm"])
    return out


def build_ref_polyiou(force=False):
    """The reference's polygon IoU (DOTA_devkit/polyiou/csrc/polyiou.cpp), compiled where it lies, behind a C shim."""
    src = os.path.join(REF, "DOTA_devkit/polyiou/csrc/polyiou.cpp")
    if not os.path.exists(src):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim_poly.cpp")
    out = os.path.join(REF_OUT, "libref_polyiou.so")
    if force or _stale(out, [shim, src]):
        _run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, shim, src])
    return out


def build_ref_shims(force=False):
    """The reference header compiled in place, four ways: {5,6}-float boxes x {CPU,CUDA} semantics."""
    if not os.path.isdir(REF):
        return []
    os.makedirs(REF_OUT, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim_iou.cpp")
    outs = []
    for tag, inc, ml in (("iou", "utils/box_iou_rotated/src", False),
                         ("nms", "utils/nms_rotated/src", False),
                         ("ml", "utils/ml_nms_rotated/src", True)):
        hdr = os.path.join(REF, inc, "box_iou_rotated_utils.h")
        for sem in ("cpu", "cudasem"):
            out = os.path.join(REF_OUT, "libref_%s_%s.so" % (tag, sem))
            outs.append(out)
            if not (force or _stale(out, [shim, hdr])):
                continue
            defs = ["-DREF_SYM(n)=ref_%s_%s_##n" % (tag, sem)] + (["-DREF_ML"] if ml else [])
            if sem == "cpu":
                _run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(REF, inc)]
                     + defs + ["-o", out, shim])
            else:
                # nvcc's host pass defines __CUDACC__, which selects the header's device-side hull sort.
                _run(["nvcc", "-O2", "-x", "cu", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
                      "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(REF, inc)]
                     + defs + ["-o", out, shim])
    return outs


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    for o in build_ref_shims(force="--force" in sys.argv):
        print(o)
    print(build_ref_polyiou(force="--force" in sys.argv))


# ---- the reference's own torch extensions, compiled from /root/reference in place ---------------

REF_EXTS_CPU = {
    # name (== TORCH_EXTENSION_NAME == the reference's module name) : sources relative to REF
    "box_iou_rotated_cuda": ["utils/box_iou_rotated/src/box_iou_rotated_cpu.cpp"],
    "nms_rotated_cuda": ["utils/nms_rotated/src/nms_rotated_cpu.cpp"],
    "ml_nms_rotated_cuda": ["utils/ml_nms_rotated/src/nms_rotated_cpu.cpp"],
}
REF_EXTS_GPU = {
    "box_iou_rotated_cuda": ["utils/box_iou_rotated/src/box_iou_rotated_cpu.cpp",
                             "utils/box_iou_rotated/src/box_iou_rotated_cuda.cu"],
    "nms_rotated_cuda": ["utils/nms_rotated/src/nms_rotated_cpu.cpp", "utils/nms_rotated/src/nms_rotated_cuda.cu"],
    "ml_nms_rotated_cuda": ["utils/ml_nms_rotated/src/nms_rotated_cpu.cpp",
                            "utils/ml_nms_rotated/src/nms_rotated_cuda.cu"],
    "deform_conv_cuda": ["models/dcn/src/deform_conv_cuda.cpp", "models/dcn/src/deform_conv_cuda_kernel.cu"],
    # needs oracle/thc_shim on the include path (THC/THC.h is gone from torch; three symbols, see the shim header)
    "orn_cuda": ["models/orn/src/vision.cpp", "models/orn/src/cpu/ActiveRotatingFilter_cpu.cpp",
                 "models/orn/src/cpu/RotationInvariantEncoding_cpu.cpp", "models/orn/src/cuda/ActiveRotatingFilter_cuda.cu",
                 "models/orn/src/cuda/RotationInvariantEncoding_cuda.cu"],
}


def build_ref_extensions(kind="cpu", names=None, verbose=False):
    """Compile the UNMODIFIED reference extension sources where they lie (torch cpp_extension,
    ninja) into oracle/_ref/ext_<kind>/<name>/<name>.so.  kind="cpu": the reference's CPU kernels
    (golden vectors, CPU baseline).  kind="gpu": its CUDA kernels for sm_100a (parity witness on
    the GPU box).  Returns {name: path}."""
    if not os.path.isdir(REF):
        return {}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils import cpp_extension
    table = REF_EXTS_CPU if kind == "cpu" else REF_EXTS_GPU
    out = {}
    for name, srcs in table.items():
        if names and name not in names:
            continue
        bdir = os.path.join(REF_OUT, "ext_%s" % kind, name)
        os.makedirs(bdir, exist_ok=True)
        so = os.path.join(bdir, name + ".so")
        full = [os.path.join(REF, s) for s in srcs]
        if _stale(so, full):
            cpp_extension.load(name=name, sources=full, build_directory=bdir, verbose=verbose,
                               extra_cflags=["-O2"] + (["-DWITH_CUDA"] if kind == "gpu" else []),
                               extra_cuda_cflags=["-O2", "-DWITH_CUDA", "-gencode", "arch=compute_100a,code=sm_100a"],
                               extra_include_paths=[os.path.join(HERE, "thc_shim")],
                               with_cuda=(kind == "gpu"), is_python_module=False)
        out[name] = so
    return out


_LOADED = {}


def load_ref_extension(name, kind="cpu"):
    """Import a prebuilt reference extension from oracle/_ref (no compilation, no /root/reference).

    The CPU-only and the CUDA build of one extension export the same (weak, inline) C++ dispatcher symbols;
    whichever is loaded first wins the dynamic symbol resolution for both.  The CUDA build's dispatcher
    serves CPU tensors too, the CPU build's does not ("Not compiled with GPU support"), so on a machine with
    a GPU the CUDA build is always loaded first."""
    import importlib.util
    key = (name, kind)
    if key in _LOADED:
        return _LOADED[key]
    so = os.path.join(REF_OUT, "ext_%s" % kind, name, name + ".so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    if kind == "cpu" and torch.cuda.is_available():
        load_ref_extension(name, "gpu")
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _LOADED[key] = mod
    return mod


# ---- the reference checkout staged for the GPU box ------------------------------------------------

STAGED = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def stage_reference_tree():
    """Copy the (read-only) reference checkout into the git-ignored baseline/_ref/ so that it travels with the gpurun
    snapshot: the -m gpu tests import the reference's UNMODIFIED Python (models/head.py, models/detector.py,
    utils/bbox_nms_rotated.py ...) from there on top of the shim modules.  Nothing under baseline/_ref is tracked,
    edited or shipped as product source.  Returns the staged root (or None where the reference is not mounted)."""
    import shutil
    if not os.path.isdir(REF):
        return STAGED if os.path.isdir(os.path.join(STAGED, "models")) else None
    os.makedirs(STAGED, exist_ok=True)
    shutil.copytree(REF, STAGED, dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "*.so", "*.o", "build"))
    return STAGED


def reference_root():
    """Where the reference's Python tree can be imported from: /root/reference in the authoring container, the
    staged copy on the GPU box; None if neither exists."""
    for root in (REF, STAGED):
        if os.path.isfile(os.path.join(root, "models", "head.py")):
            return root
    return None
