"""Build libcreste_torch_ops.so (the C++ registration of the `creste::` ops) in-tree with g++ against the installed
libtorch: python creste_public_b200/csrc_torch/build.py"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libcreste_torch_ops.so")


def build(verbose=False):
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(HERE, "creste_torch_ops.cpp")
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(HERE, "..", "..", "include", "creste_b200.h"))):
        return OUT
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    csrc = os.path.abspath(os.path.join(HERE, "..", "csrc"))
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", src, "-o", OUT] + inc + [
        f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda", f"-L{csrc}", "-lcreste_b200",
        f"-Wl,-rpath,{libdir}", "-Wl,-rpath,$ORIGIN/../csrc", "-Wl,--no-as-needed"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libcreste_torch_ops.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
