"""Compile the reference's own KiloNeRF kernels (the a9 path) into oracle/_ref/kilonerf_ref*.so.

    python oracle/build_kilo_ref.py

Test infrastructure: the resulting module is imported only by tests/test_gpu_kilo_ref.py (and skipped when
absent).  Sources are compiled WHERE THEY LIE under /root/reference/cuda -- generate_inputs.cu,
network_eval.cu, integrate.cu, utils.cu -- for sm_100a with plain nvcc and the torch headers of this image;
the MAGMA / OpenGL parts of the extension (multimatmul.cu, render_to_screen.cpp, cuda/setup.py:7,18-19) are
not needed for this path and are left out, so the reference's own setup.py is not used.  Outputs go to
oracle/_ref/ (git-ignored, travels to the GPU box).  Needs /root/reference; a no-op elsewhere.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/cuda"
OUT = os.path.join(HERE, "_ref")
NAME = "kilonerf_ref"
SOURCES = ["generate_inputs.cu", "network_eval.cu", "integrate.cu", "utils.cu"]


def build(force=False):
    if not os.path.isdir(REF):
        print("build_kilo_ref: %s not present, nothing to do" % REF)
        return None
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, NAME + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and not force:
        return so
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = ["-I" + p for p in ce.include_paths("cuda")] + ["-I" + sysconfig.get_paths()["include"], "-I" + REF]
    common = ["-O3", "-std=c++17", "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT, src[:-3] + ".o")
        objs.append(obj)
        if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(os.path.join(REF, src)) and not force:
            continue    # left over from an interrupted build (each torch-header translation unit takes minutes)
        # utils.cu uses fprintf without including <stdio.h> (cuda/utils.cu:14): pre-include it, sources untouched
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
               "--pre-include", "stdio.h", "-w"] + common + inc + ["-c", os.path.join(REF, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bobj = os.path.join(OUT, "binding.o")
    objs.append(bobj)
    if not os.path.exists(bobj) or force:
        cmd = ["g++", "-fPIC", "-w"] + common + inc + ["-c", os.path.join(HERE, "kilo_ref_binding.cpp"), "-o", bobj]
        procs.append(("binding", subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("build_kilo_ref: %s failed:\n%s" % (name, out[-4000:]))
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-shared", "-o", so] + objs + ["-L" + libdir, "-Wl,-rpath," + libdir, "-lc10", "-ltorch", "-ltorch_cpu",
                                                  "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                                                  "-L/usr/local/cuda/lib64", "-lcudart"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("build_kilo_ref: link failed:\n" + p.stderr[-4000:])
    for o in objs:
        os.remove(o)
    return so


if __name__ == "__main__":
    print("built", build(force="--force" in sys.argv))
