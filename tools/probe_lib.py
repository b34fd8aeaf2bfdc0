"""Build / load tools/csrc/fd_probe.cu as its own shared library (development probes are not part of the
product library libflowdec_b200.so)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load():
    out = os.path.join(HERE, "libfd_probe.so")
    src = os.path.join(HERE, "csrc", "fd_probe.cu")
    api = os.path.join(ROOT, "flowdec_b200", "csrc", "fd_api.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode",
                               "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler",
                               "-fPIC", "-shared", "-I", os.path.join(ROOT, "flowdec_b200", "csrc"), src, api, "-o",
                               out, "-lcudart"])
    return ctypes.CDLL(out)
