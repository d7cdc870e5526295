"""Round-2 experiment: epilogue variants of the CTA-pair persistent GEMM (tools/probe/conv_pair_probe2.cu) on the head's
1x1 layer shapes, with and without the residual.  Wrap in `timeout`: a protocol mistake deadlocks.

    gpurun --timeout 400 -- 'timeout -s KILL 250 python tools/probe/run_conv_pair_probe2.py'
"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "conv_pair2.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                       os.path.join(here, "conv_pair_probe2.cu"), "-o", so, "-lcudart"])
lib = ctypes.CDLL(so)
P, I = ctypes.c_void_p, ctypes.c_int
g = torch.Generator().manual_seed(0)
def run(M, N, K, BN, has_res, epi, knock=0, stages=0, check=True):
    x = torch.randn(M, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
    sc = (torch.rand(N, generator=g) + 0.5).cuda()
    sh = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).half().cuda() if has_res else None
    y = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    ms = ctypes.c_float(0)
    rc = lib.conv_pair2_run(P(x.data_ptr()), I(K), P(w.data_ptr()), P(y.data_ptr()), I(N), P(res.data_ptr() if has_res else 0),
                            I(N), I(M), I(N), I(K), I(BN), I(1), P(sc.data_ptr()), P(sh.data_ptr()), I(epi), I(20), ctypes.byref(ms),
                            I(knock), I(stages))
    err = -1.0
    if check and knock == 0:
        ref = x.float() @ w.float().t() * sc + sh
        if has_res:
            ref = ref + res.float()
        ref = torch.relu(ref)
        err = float((y.float() - ref).abs().max()) / float(ref.abs().max())
    return rc, err, ms.value * 1e3


for (M, N, K, BN, has_res) in [(34496, 1024, 256, 256, False), (34496, 1024, 256, 256, True), (34496, 1024, 1088, 256, False),
                              (34496, 256, 1024, 256, False), (25088, 256, 480, 256, False)]:
    for epi in (0, 2, 4, 5, 6):
        line = "M=%d N=%d K=%d res=%d epi%d |" % (M, N, K, has_res, epi)
        for knock in (0, 4, 5):
            rc, err, us = run(M, N, K, BN, has_res, epi, knock)
            line += " knock%d rc=%d err=%.0e %6.1f us |" % (knock, rc, err, us)
        print(line, flush=True)
