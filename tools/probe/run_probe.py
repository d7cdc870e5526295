import ctypes, os, subprocess, sys
import numpy as np, torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "probe.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       os.path.join(here, "umma_shift_probe.cu"), "-o", so, "-lcudart"])
lib = ctypes.CDLL(so)
P = 256
g = torch.Generator().manual_seed(0)
x = torch.randn(P, 64, generator=g).half().cuda()
w = torch.randn(64, 64, generator=g).half().cuda()
out = torch.zeros(128, 64, device="cuda")
for S in (8, 10, 11, 9):
    for r0 in (0, 1, 3, 8, 13):
        for bo in (0, 1):
            if r0 + 15 * S + 8 > P: continue
            rc = lib.probe_run(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), P, r0, S, bo, ctypes.c_void_p(out.data_ptr()))
            rows = torch.tensor([r0 + gi * S + i for gi in range(16) for i in range(8)], device="cuda")
            ref = x[rows].float() @ w.float().t()
            err = float((out - ref).abs().max())
            print("S=%2d r0=%2d base_offset=%d rc=%d max_err=%.4f %s" % (S, r0, bo, rc, err, "OK" if err < 0.05 else "MISMATCH"))
