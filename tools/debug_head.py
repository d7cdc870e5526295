"""GPU debug: compare every stage of TwoBranchNet.forward_act against the oracle (torch CPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import step_b200
from step_b200 import synth, engine as E, _lib as L
from step_b200.engine import Act
from step_b200.two_branch import _packed, conv2d
from oracle import model as om, tubes as ot

fp16 = len(sys.argv) > 1 and sys.argv[1] == "fp16"
cfg = synth.make_cfg(fp16=fp16, T=2, max_iter=1, NUM_CHUNKS={1: 1}, image_size=(112, 112))
sd_b, sd_h = synth.base_net_state_dict(), synth.head_state_dict(100, cfg)
base = step_b200.BaseNet(cfg); base.load_state_dict(sd_b); base = base.cuda().eval()
head = step_b200.TwoBranchNet(cfg); head.load_state_dict(sd_h); head = head.cuda().eval(); head.set_device("cuda:0")
roi = step_b200.ROINet("align", 7)
x = synth.make_clips(1, 8, 112, 112)
tubes = synth.make_proposals(1, 3, 2, 112, 112)
def rel(name, got, ref):
    got = got.float().cpu(); ref = ref.float()
    print("%-14s shape %-22s max|ref| %9.4f  maxerr %9.5f  rel %.2e" % (name, tuple(ref.shape), ref.abs().max(), (got-ref).abs().max(), (got-ref).abs().max()/ref.abs().max()))
with torch.no_grad():
    feat = base.forward_act(x.cuda())
    cf_ref = om.base_net(x, sd_b)
    rel("conv_feat", feat.logical(), cf_ref)
    flat_np, nums = ot.flatten_tubes(tubes, True)
    flat = torch.from_numpy(flat_np).cuda()
    R, T = flat.shape[0], flat.shape[1]
    code = feat.code
    cat = Act.empty(R, T, 7, 7, 1088, code, "cuda")
    roi.pool_into(feat, flat, cat.frames().slice(0, 832), T, feat.T, 0)
    pooled_ref = om.roi_net(cf_ref, torch.from_numpy(flat_np)).view(R, T, 832, 7, 7)
    rel("pooled", cat.slice(0, 832).logical(), pooled_ref)
    g = head.i3d_conv[0](cat.slice(0, 832)); gr = om.mixed(pooled_ref.permute(0, 2, 1, 3, 4), sd_h, "i3d_conv.0.")
    rel("mixed_5b", g.logical(), gr.permute(0, 2, 1, 3, 4))
    g = head.i3d_conv[1](g); gr = om.mixed(gr, sd_h, "i3d_conv.1.")
    rel("mixed_5c", g.logical(), gr.permute(0, 2, 1, 3, 4))
    w, b = _packed(head.downsample, code)
    gconv = cat.slice(832, 256)
    E.conv(g, w, None, b, gconv, (1, 1, 1), relu=False)
    gcr = F.conv3d(gr, sd_h["downsample.weight"], sd_h["downsample.bias"])
    rel("downsample", gconv.logical(), gcr.permute(0, 2, 1, 3, 4))
    lfr = torch.cat([pooled_ref.permute(0, 2, 1, 3, 4), gcr], 1).permute(0, 2, 1, 3, 4).contiguous().view(R * T, 1088, 7, 7)
    fr = cat.frames()
    rel("cat", fr.logical()[:, 0], lfr)
    m = head.local_conv[0]
    res = conv2d(m.conv1, fr, False); rel("b0.conv1", res.logical()[:, 0], F.conv2d(lfr, sd_h["local_conv.0.conv1.weight"]))
    o = conv2d(m.conv2, fr, True); orf = F.relu(F.conv2d(lfr, sd_h["local_conv.0.conv2.weight"])); rel("b0.conv2", o.logical()[:, 0], orf)
    o = conv2d(m.conv3, o, True); orf = F.relu(F.conv2d(orf, sd_h["local_conv.0.conv3.weight"], padding=1)); rel("b0.conv3", o.logical()[:, 0], orf)
    lf = fr
    lr = lfr
    for i, rs in enumerate([True, False, False]):
        lf = head.local_conv[i](lf); lr = om._bottleneck(lr, sd_h, "local_conv.%d." % i, rs)
        rel("bottleneck%d" % i, lf.logical()[:, 0], lr)
    w2, b2 = _packed(head.downsample2, code)
    lf2 = Act.empty(R * T, 1, 7, 7, 256, code, "cuda")
    E.conv(lf, w2, None, b2, lf2, (1, 1, 1), relu=False)
    l2r = F.conv2d(lr, sd_h["downsample2.weight"], sd_h["downsample2.bias"])
    rel("downsample2", lf2.logical()[:, 0], l2r)
    wr, br = head._reg_weight("local_reg", code)
    loc = E.linear_small_n(lf2.buf, R * T, 12544, 12544, wr, br, 4)
    locr = F.linear(l2r.reshape(R * T, -1), sd_h["local_reg.weight"], sd_h["local_reg.bias"])
    rel("local_reg", loc, locr)
    print(loc.cpu()[:3]); print(locr[:3])
    prob, loc2, first, last = head.forward_act(cat, None, None)
    pr, lr_, fr_, la_ = om.two_branch(pooled_ref, sd_h, cfg.T)
    rel("prob", prob, pr); rel("loc(full)", loc2, lr_); rel("first", first, fr_); rel("last", last, la_)
