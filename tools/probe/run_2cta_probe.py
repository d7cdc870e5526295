"""Round-2 starting point: builds and runs tools/probe/umma_2cta_probe.cu on a GPU box (wrap in `timeout`: a protocol
mistake in a 2-CTA kernel deadlocks rather than faults).

    gpurun --timeout 300 -- 'timeout 120 python tools/probe/run_2cta_probe.py'
"""
import ctypes, os, subprocess
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "probe2.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       os.path.join(here, "umma_2cta_probe.cu"), "-o", so, "-lcudart"])
lib = ctypes.CDLL(so)
g = torch.Generator().manual_seed(0)
for N, K in ((128, 64), (256, 256), (192, 128)):
    a = torch.randn(256, K, generator=g).half().cuda()
    b = torch.randn(N, K, generator=g).half().cuda()
    out = torch.zeros(256, N, device="cuda")
    rc = lib.probe2_run(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), N, K, ctypes.c_void_p(out.data_ptr()))
    ref = a.float() @ b.float().t()
    err = float((out - ref).abs().max())
    print("N=%3d K=%3d rc=%d max_err=%.4f %s" % (N, K, rc, err, "OK" if rc == 0 and err < 0.05 * K ** 0.5 else "MISMATCH"))
