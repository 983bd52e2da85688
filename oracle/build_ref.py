"""Compile the reference's own CPU NMS, unmodified, into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Source compiled where it lies: /root/reference/mmdet/ops/nms/src/nms_cpu.cpp (never copied
into this repo).  Output: oracle/_ref/nms_cpu_ref.so, a pybind11 module exposing the
reference's `nms(dets, thr)` and `soft_nms(dets, thr, method, sigma, min_score)`.
oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot.

The reference's CUDA sources (nms_kernel.cu, roi_align_kernel*.cu) do NOT build against
torch 2.11 (THC/THC.h and AT_CHECK are gone) and are treated as unbuildable; see DESIGN.md.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/mmdet/ops/nms/src/nms_cpu.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "nms_cpu_ref.so")


def build(force=False):
    if not os.path.exists(SRC):
        return os.path.exists(OUT)          # GPU box: use the prebuilt file, if any
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return True
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = cpp_extension.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w",
           "-DTORCH_EXTENSION_NAME=nms_cpu_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in inc]
    cmd += [SRC, "-o", OUT, "-L" + libdir, "-Wl,-rpath," + libdir,
            "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python"]
    subprocess.check_call(cmd)
    return True


def load():
    """Import the compiled reference module, or return None if it was never built."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("nms_cpu_ref", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference sources not present; nothing built")
