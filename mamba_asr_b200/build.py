"""In-tree build of libconmamba_b200.so (sm_100a only) with nvcc.

    python -m mamba_asr_b200.build [--force]

One object per .cu (compiled in parallel), linked into ``mamba_asr_b200/lib/libconmamba_b200.so``.  The
shared object is git-ignored but travels to the GPU box with the repo snapshot.  A content hash of the
sources + flags skips up-to-date objects.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJDIR = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libconmamba_b200.so")

SOURCES = ["scan_fwd.cu", "scan_fwd_cl.cu", "scan_fwd_sp.cu", "scan_fwd_lc.cu", "scan_fwd_wg.cu", "scan_bwd.cu", "scan_bwd_cl.cu", "scan_bwd_sp.cu", "scan_bwd_lc.cu", "scan_bwd_wg.cu", "conv.cu", "fbank.cu", "fbank_dft.cu", "layernorm.cu", "dwconv.cu", "colsum.cu", "step.cu", "fused_ln.cu", "act.cu", "tsmm.cu", "ln_act.cu", "optim.cu", "ctc.cu", "stem.cu", "reduce.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda",
    "-I" + INCLUDE, "-I" + CSRC,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the ConMamba B200 kernels cannot be built")


def _digest(paths, extra):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return hs


def build(force=False, verbose=False, variant=None, extra_flags=None):
    """variant / extra_flags: tuning experiments - a second library lib/libconmamba_b200_<variant>.so compiled with extra
    -D flags next to the default one (select it at run time with CM_LIB_PATH); never used by the product path."""
    objdir = OBJDIR if variant is None else OBJDIR + "_" + variant
    lib = LIB if variant is None else LIB.replace(".so", "_%s.so" % variant)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = _headers()
    extra = os.environ.get("CM_NVCC_EXTRA", "").split() + list(extra_flags or [])     # e.g. -DCM_FWD_MINB=8
    flags = " ".join(NVCC_FLAGS + extra)

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dg = _digest([path] + headers, flags)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
            return obj, False
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)
        return obj, True

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(lib):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    ext = sys.argv[sys.argv.index("--extra") + 1].split() if "--extra" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var, extra_flags=ext))
