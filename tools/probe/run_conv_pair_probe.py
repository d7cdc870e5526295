"""Round-2 experiment: the CTA-pair (cta_group::2) persistent GEMM of tools/probe/conv_pair_probe.cu against torch and
against the in-library persistent kernel on the head's 1x1 layer shapes.  NOT YET RUN (written after the round-1 GPU
budget was spent); wrap in `timeout`, a protocol mistake deadlocks.

    gpurun --timeout 400 -- 'python -m step_b200.build >/dev/null; timeout -s KILL 200 python tools/probe/run_conv_pair_probe.py'
"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "conv_pair.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                       os.path.join(here, "conv_pair_probe.cu"), "-o", so, "-lcudart"])
lib = ctypes.CDLL(so)
P, I = ctypes.c_void_p, ctypes.c_int
g = torch.Generator().manual_seed(0)
cases = [(300, 128, 64, 128), (1000, 256, 192, 256), (34496, 1024, 1088, 256), (34496, 256, 1024, 256), (34496, 1024, 256, 256)]
for M, N, K, BN in cases:
    x = torch.randn(M, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
    sc = (torch.rand(N, generator=g) + 0.5).cuda()
    sh = torch.randn(N, generator=g).cuda()
    y = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    ms = ctypes.c_float(0)
    rc = lib.conv_pair_run(P(x.data_ptr()), I(K), P(w.data_ptr()), P(y.data_ptr()), I(N), I(M), I(N), I(K), I(BN), I(1),
                           P(sc.data_ptr()), P(sh.data_ptr()), I(20), ctypes.byref(ms))
    ref = torch.relu(x.float() @ w.float().t() * sc + sh)
    err = float((y.float() - ref).abs().max()) / float(ref.abs().max())
    tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value else 0.0
    print("M=%6d N=%4d K=%4d BN=%3d rc=%d rel_err=%.2e  %.1f us  %.0f TFLOP/s" % (M, N, K, BN, rc, err, ms.value * 1e3, tf))
try:   # the in-library kernel on the same shapes, for comparison
    from step_b200 import engine as E, _lib as L
    from step_b200.engine import Act
    for M, N, K, BN in cases[2:]:
        x = Act(torch.randn(M // 49, 1, 7, 7, K, generator=g).half().cuda())
        wq = (torch.randn(N, 1, K, generator=g) / K ** 0.5).half().cuda()
        out = Act(torch.empty(M // 49, 1, 7, 7, N, dtype=torch.float16, device="cuda"))
        one, zero = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
        f = lambda: E.conv(x, wq, one, zero, out, (1, 1, 1))
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        print("library persistent kernel  M=%6d N=%4d K=%4d: %.1f us" % (M, N, K, e0.elapsed_time(e1) / 20 * 1e3))
except Exception as ex:
    print("library comparison skipped:", ex)
