"""Diagnostic: per-parameter error of the device head backward vs tests/golden/head_grads.npz for several loss scales."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import step_b200
from step_b200 import synth, training
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "head_grads.npz"))
T_, chunks, n, _ = synth.LOSS_CASES["c1"]
cfg = synth.make_cfg(fp16=True, T=T_, max_iter=1, NUM_CHUNKS={1: chunks}, image_size=(112, 112))
_, _, feat, tb, tg = synth.make_loss_case("c1", cfg.num_classes)
net = step_b200.TwoBranchNet(cfg); net.load_state_dict(synth.head_state_dict(100, cfg)); net = net.cuda().eval(); net.set_device("cuda:0")
names = {p: k for k, p in net.named_parameters()}
for S in (64.0, 1024.0, 16384.0, 262144.0):
    r = training.head_forward_backward(net, feat.cuda(), tb.cuda(), tg.cuda(), loss_scale=S)
    got = {names[p]: v for p, v in r["grads"].items()}
    rows = []
    for key in g.files:
        if not key.startswith("gn:"):
            continue
        k = key[3:]
        if k not in got:
            continue
        ref_n = float(g[key][0]); gn = float(got[k].double().norm())
        lead = got[k].reshape(-1)[:8].cpu().numpy()
        le = float(np.abs(lead - g["gh:" + k]).max() / max(np.abs(g["gh:" + k]).max(), ref_n / got[k].numel() ** 0.5))
        rows.append((abs(gn - ref_n) / ref_n, le, k))
    rows.sort(reverse=True)
    fg = r["feat_grad"]
    print("S=%g  worst norm rel %.3e (%s)  worst lead %.3e (%s)  feat_grad rel %.3e  finite %s" % (
        S, rows[0][0], rows[0][2], max(r_[1] for r_ in rows), max(rows, key=lambda r_: r_[1])[2],
        abs(float(fg.double().norm()) - float(g["feat_grad_norm"][0])) / float(g["feat_grad_norm"][0]), bool(torch.isfinite(fg).all())))
    for r_ in sorted(rows, key=lambda r_: -r_[1])[:5]:
        print("     lead err %.3e norm err %.3e  %s" % (r_[1], r_[0], r_[2]))
