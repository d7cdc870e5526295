"""TEST INFRASTRUCTURE ONLY -- builds the oracle artefacts.  Nothing under step_b200/ imports this.

1. ``oracle/_build/libstep_oracle.so``  <- oracle/step_oracle.c   (plain C restatement, gcc)
2. ``oracle/_ref/_C.so``                <- the reference's *own* CPU C++ ops, compiled from the
   sources where they lie (``/root/reference/external/maskrcnn_benchmark/csrc/{vision.cpp,cpu/*.cpp}``)
   with g++ directly -- the reference's setup.py/build system is not run, no source is copied.
   Only possible where /root/reference exists (the build container); the GPU box uses the prebuilt .so
   that travels with the snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/external/maskrcnn_benchmark/csrc"


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_c_oracle(verbose=False):
    src = os.path.join(HERE, "step_oracle.c")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libstep_oracle.so")
    if _newer(out, [src]):
        return out
    # -ffp-contract=off: the oracle must round every fp32 op like the reference's x86 build (no FMA fusion)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", out, src, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


def build_reference_ops(verbose=False):
    """Compile the reference's CPU `_C` extension into oracle/_ref/_C.so.  Returns path or None."""
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "_C.so")
    if not os.path.isdir(REF_CSRC):
        return out if os.path.exists(out) else None
    srcs = [os.path.join(REF_CSRC, "vision.cpp"),
            os.path.join(REF_CSRC, "cpu", "ROIAlign_cpu.cpp"),
            os.path.join(REF_CSRC, "cpu", "nms_cpu.cpp")]
    shim = os.path.join(HERE, "ref_shim.h")
    if _newer(out, srcs + [shim]):
        return out
    os.makedirs(out_dir, exist_ok=True)
    import torch
    from torch.utils import cpp_extension
    incs = cpp_extension.include_paths() + [sysconfig.get_paths()["include"], REF_CSRC]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
           "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-include", shim]
    for i in incs:
        cmd += ["-isystem", i]
    cmd += srcs + ["-o", out, "-L" + torch_lib, "-Wl,-rpath," + torch_lib,
                   "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_c_oracle(verbose=True))
    print(build_reference_ops(verbose=True))
