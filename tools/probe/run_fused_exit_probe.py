"""Round-2 experiment: fused bottleneck exit (conv3 + residual + ReLU -> next conv1 / downsample2) on CTA pairs
(tools/probe/fused_exit_probe.cu) against torch and against the two library launches it replaces.
Wrap in `timeout`: a protocol mistake deadlocks.

    gpurun --timeout 400 -- 'timeout -s KILL 200 python tools/probe/run_fused_exit_probe.py'
"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "fused_exit.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                       os.path.join(here, "fused_exit_probe.cu"), "-o", so, "-lcudart"])
lib = ctypes.CDLL(so)
P, I = ctypes.c_void_p, ctypes.c_int
g = torch.Generator().manual_seed(0)
for M, store_y, relu2, bias in [(256, 1, 1, 0), (1000, 1, 1, 0), (34496, 1, 1, 0), (34496, 0, 0, 1)]:
    h = torch.randn(M, 256, generator=g).half().cuda()
    w3 = (torch.randn(1024, 256, generator=g) / 16).half().cuda()
    x = torch.randn(M, 1024, generator=g).half().cuda()
    w1 = (torch.randn(256, 1024, generator=g) / 32).half().cuda()
    sh = torch.randn(256, generator=g).cuda() if bias else None
    y = torch.zeros(M, 1024, dtype=torch.float16, device="cuda") if store_y else None
    z = torch.zeros(M, 256, dtype=torch.float16, device="cuda")
    ms = ctypes.c_float(0)
    rc = lib.fused_exit_run(P(h.data_ptr()), I(256), P(w3.data_ptr()), P(x.data_ptr()), I(1024), P(w1.data_ptr()),
                            P(y.data_ptr() if store_y else 0), I(1024), P(z.data_ptr()), I(256), I(M), P(sh.data_ptr() if bias else 0),
                            I(relu2), I(20), ctypes.byref(ms))
    yref = torch.relu(h.float() @ w3.float().t() + x.float())
    zref = yref.half().float() @ w1.float().t()
    if bias:
        zref = zref + sh
    if relu2:
        zref = torch.relu(zref)
    ey = float((y.float() - yref).abs().max() / yref.abs().max()) if store_y else -1.0
    ez = float((z.float() - zref).abs().max() / zref.abs().max())
    print("M=%6d store_y=%d relu2=%d bias=%d rc=%d  err_y %.2e err_z %.2e  %.1f us" % (M, store_y, relu2, bias, rc, ey, ez, ms.value * 1e3), flush=True)
# the two library launches it replaces, for comparison
try:
    from step_b200 import engine as E
    from step_b200.engine import Act
    M = 34496
    hA = Act(torch.randn(M // 49, 1, 7, 7, 256, generator=g).half().cuda())
    xA = Act(torch.randn(M // 49, 1, 7, 7, 1024, generator=g).half().cuda())
    w3p = (torch.randn(1024, 1, 256, generator=g) / 16).half().cuda()
    w1p = (torch.randn(256, 1, 1024, generator=g) / 32).half().cuda()
    yA = Act(torch.empty(M // 49, 1, 7, 7, 1024, dtype=torch.float16, device="cuda"))
    zA = Act(torch.empty(M // 49, 1, 7, 7, 256, dtype=torch.float16, device="cuda"))
    def two():
        E.conv(hA, w3p, None, None, yA, (1, 1, 1), relu=True, residual=xA)
        E.conv(yA, w1p, None, None, zA, (1, 1, 1), relu=True)
    for _ in range(3): two()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): two()
    e1.record(); torch.cuda.synchronize()
    print("library: conv3+res+relu then conv1: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
except Exception as ex:
    print("library comparison skipped:", ex)
