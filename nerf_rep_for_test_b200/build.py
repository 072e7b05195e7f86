"""Build libnerfb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m nerf_rep_for_test_b200.build [--force]

Each csrc/*.cu is compiled to an object (in parallel) and linked into
nerf_rep_for_test_b200/libnerfb200.so with the CUDA runtime linked statically,
so the library has no torch and no libcudart.so dependency.  Objects are
rebuilt only when a source or header is newer.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiments: NERFB200_VARIANT=<tag> NERFB200_NVCC_FLAGS="-D..." builds libnerfb200_<tag>.so next to the product
# library (load it with NERFB200_LIB=<path>); the default build is the product
_VAR = os.environ.get("NERFB200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "libnerfb200%s.so" % ("_" + _VAR if _VAR else ""))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libnerfb200.so cannot be built")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "nerfb200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, log):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(spath), _newest_header())):
        return obj, ""
    cmd = [nvcc()] + ARCH + FLAGS + os.environ.get("NERFB200_NVCC_FLAGS", "").split() + ["-c", spath, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), sources()))
    objs = [r[0] for r in results]
    log = "\n".join(r[1] for r in results if r[1])
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            print(log)
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [nvcc()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print("built", path)
