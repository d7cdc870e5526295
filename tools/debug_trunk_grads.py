"""Diagnostic: per-layer error of the device trunk backward vs the oracle's torch-CPU autograd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import step_b200
from step_b200 import synth, training
from oracle import model as om
S = float(sys.argv[1]) if len(sys.argv) > 1 else 1024.0
cfg = synth.make_cfg(fp16=True, T=2, max_iter=1, NUM_CHUNKS={1: 1}, image_size=(64, 64))
net = step_b200.BaseNet(cfg); net.load_state_dict(synth.base_net_state_dict()); net = net.cuda().eval()
x = synth.make_clips(1, 8, 64, 64, seed=4321)
proj = torch.randn((1, 2, 832, 4, 4), generator=torch.Generator().manual_seed(99))
feat, grads = training.trunk_forward_backward(net, x.cuda(), lambda f: (proj / proj.numel()).permute(0, 1, 3, 4, 2).contiguous().cuda(), loss_scale=S)
names = {p: k for k, p in net.named_parameters()}
got = {names[p]: v for p, v in grads.items()}
sd = {k: v.clone().requires_grad_(k.endswith("conv3d.weight")) for k, v in synth.base_net_state_dict().items()}
cf = om.base_net(x.clone(), sd)
((cf * proj).sum() / cf.numel()).backward()
print("feat rel err", float((feat.logical().float().cpu() - cf.detach()).abs().max() / cf.abs().max()))
for k in sorted(got, key=lambda s: (int(s.split(".")[1]), s)):
    ref = sd[k].grad
    rn, gn = float(ref.double().norm()), float(got[k].double().norm())
    rel = float((got[k].cpu().double() - ref.double()).norm() / ref.double().norm())
    print("%-44s shape %-22s norm ref %.4e got %.4e (%+.1f%%)  relL2 %.3f" % (k, tuple(ref.shape), rn, gn, 100 * (gn - rn) / rn, rel))
