"""Diagnostic: thin stride-1 3x3x3 convolutions (few channels, small maps, residual accumulate) vs torch fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from step_b200 import engine as E
from step_b200.engine import Act
torch.manual_seed(0)
for (N, T, H, W, Cin, Cout, k, res) in [(1, 4, 8, 8, 32, 16, (3, 3, 3), True), (1, 4, 8, 8, 32, 16, (3, 3, 3), False), (1, 4, 8, 8, 16, 32, (3, 3, 3), False),
                                        (1, 4, 8, 8, 176, 192, (1, 1, 1), True), (1, 2, 4, 4, 48, 16, (3, 3, 3), True), (1, 4, 8, 8, 96, 32, (3, 3, 3), True)]:
    x = Act(torch.randn(N, T, H, W, Cin, device="cuda").half())
    taps = k[0] * k[1] * k[2]
    w = (torch.randn(Cout, taps, Cin, device="cuda") / (Cin * taps) ** 0.5).half()
    r0 = torch.randn(N, T, H, W, Cout, device="cuda").half()
    out = Act(r0.clone())
    E.conv(x, w, None, None, out, k, (1, 1, 1), tuple(kk // 2 for kk in k), relu=False, residual=out if res else None, out_dims=(T, H, W))
    xs = x.buf.float().permute(0, 4, 1, 2, 3)
    ws = w.float().view(Cout, k[0], k[1], k[2], Cin).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(xs, ws, padding=tuple(kk // 2 for kk in k)).permute(0, 2, 3, 4, 1)
    if res:
        ref = ref + r0.float()
    err = float((out.buf.float() - ref).abs().max() / ref.abs().max())
    nrm = float(out.buf.float().norm() / ref.norm())
    print("N%d T%d %dx%d Cin %3d Cout %3d k%s res=%d: max rel err %.2e, norm ratio %.4f" % (N, T, H, W, Cin, Cout, k, res, err, nrm))
