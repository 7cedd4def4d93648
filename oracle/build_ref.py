"""Compile the reference's only native component, UNMODIFIED, from where it
lies (/root/reference/libs/utils/csrc/nms_cpu.cpp) into oracle/_ref/.

TEST INFRASTRUCTURE. Output: oracle/_ref/nms_1d_cpu.so (a CPython extension
module, same name the reference's libs/utils/setup.py:11 gives it). Plain g++
with torch's include paths — the reference's own build system is not run.
oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AVDF_REFERENCE_ROOT", "/root/reference") + "/libs/utils/csrc/nms_cpu.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "nms_1d_cpu.so")


def build(force=False):
    if not os.path.isfile(SRC):
        return None
    if os.path.isfile(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fopenmp",
           "-DTORCH_EXTENSION_NAME=nms_1d_cpu", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           *inc, SRC, "-o", OUT, "-L" + libdir, "-Wl,-rpath," + libdir,
           "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
