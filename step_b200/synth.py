"""Synthetic, seed-deterministic inputs for the STEP hot path.

There is no dataset or checkpoint offline, so parity tests and the benchmark use:
  * state dicts with the reference's key names (networks.py:107-132, two_branch.py:164-203) --
    He-normal conv weights, randomised BatchNorm statistics (the default init collapses
    activations to ~2e-5 and makes tolerances meaningless);
  * clips ~ clamp(randn, -1, 1)  (the reference's scale_norm=2 input range, augmentations.py:80-82);
  * N grid proposals per clip built like data/data_utils.py:19-45 plus a centre and a full box.

Everything here is host-side numpy/torch-CPU generation; no compute on the hot path.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

# (in_channels, [b0, b1a, b1b, b2a, b2b, b3]) -- i3dpt.py:213-231
MIXED_PLAN = {
    "3b": (192, [64, 96, 128, 16, 32, 32]), "3c": (256, [128, 128, 192, 32, 96, 64]),
    "4b": (480, [192, 96, 208, 16, 48, 64]), "4c": (512, [160, 112, 224, 24, 64, 64]),
    "4d": (512, [128, 128, 256, 24, 64, 64]), "4e": (512, [112, 144, 288, 32, 64, 64]),
    "4f": (528, [256, 160, 320, 32, 128, 128]),
    "5b": (832, [256, 160, 320, 32, 128, 128]), "5c": (832, [384, 192, 384, 48, 128, 128]),
}
# position in BaseNet.base_model (nn.Sequential, networks.py:120-132)
TRUNK_MIXED = {5: "3b", 6: "3c", 8: "4b", 9: "4c", 10: "4d", 11: "4e", 12: "4f"}


def make_cfg(**kw):
    """The cfg attributes the ported modules read, C4 defaults."""
    d = dict(base_net="i3d", kinetics_pretrain=None, freeze_stats=True, freeze_affine=True, fp16=False,
             num_classes=60, T=8, fc_dim=256, dropout=0.3, pool_size=7, pool_mode="align",
             no_context=True, max_iter=3, temporal_mode="predict", NUM_CHUNKS={1: 1, 2: 1, 3: 1},
             image_size=(224, 224))
    d.update(kw)
    return SimpleNamespace(**d)


def _conv(g, sd, key, cout, cin, k, bias=False, std=None):
    fan_in = cin * int(np.prod(k))
    std = math.sqrt(2.0 / fan_in) if std is None else std
    sd[key + ".weight"] = torch.randn((cout, cin) + tuple(k), generator=g) * std
    if bias:
        sd[key + ".bias"] = torch.randn(cout, generator=g) * 0.1


def _bn(g, sd, key, c):
    sd[key + ".weight"] = torch.rand(c, generator=g) * 0.4 + 0.8
    sd[key + ".bias"] = torch.randn(c, generator=g) * 0.1
    sd[key + ".running_mean"] = torch.randn(c, generator=g) * 0.1
    sd[key + ".running_var"] = torch.rand(c, generator=g) + 0.5
    sd[key + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _unit(g, sd, p, cin, cout, k):
    _conv(g, sd, p + "conv3d", cout, cin, k)
    _bn(g, sd, p + "batch3d", cout)


def _mixed(g, sd, p, name):
    cin, o = MIXED_PLAN[name]
    _unit(g, sd, p + "branch_0.", cin, o[0], (1, 1, 1))
    _unit(g, sd, p + "branch_1.0.", cin, o[1], (1, 1, 1))
    _unit(g, sd, p + "branch_1.1.", o[1], o[2], (3, 3, 3))
    _unit(g, sd, p + "branch_2.0.", cin, o[3], (1, 1, 1))
    _unit(g, sd, p + "branch_2.1.", o[3], o[4], (3, 3, 3))
    _unit(g, sd, p + "branch_3.1.", cin, o[5], (1, 1, 1))


def base_net_state_dict(seed=1234):
    """Keys of BaseNet.state_dict() (networks.py:50-67, 120-132): base_model.{0..12}.*"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _unit(g, sd, "base_model.0.", 3, 64, (7, 7, 7))
    _unit(g, sd, "base_model.2.", 64, 64, (1, 1, 1))
    _unit(g, sd, "base_model.3.", 64, 192, (3, 3, 3))
    for idx, name in TRUNK_MIXED.items():
        _mixed(g, sd, "base_model.%d." % idx, name)
    return sd


def context_net_state_dict(seed=4321):
    """Keys of ContextNet.state_dict() (two_branch.py:113-130): i3d_conv_context.{1,2}.*"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _mixed(g, sd, "i3d_conv_context.1.", "5b")
    _mixed(g, sd, "i3d_conv_context.2.", "5c")
    return sd


def head_state_dict(seed, cfg, reg_std=5e-5, cls_std=2e-3):
    """Keys of TwoBranchNet.state_dict() (two_branch.py:164-203).

    The regressor / classifier weights get a small fixed std instead of the reference's
    xavier-normal init (two_branch.py:344-353): with He-scaled features xavier makes every
    delta O(10) -> exp() saturates -> all tubes collapse to the whole-image box in
    valid_tubes, which would make the progressive loop a degenerate test."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _mixed(g, sd, "i3d_conv.0.", "5b")
    _mixed(g, sd, "i3d_conv.1.", "5c")
    fc, ps = cfg.fc_dim, cfg.pool_size
    D = fc * ps * ps
    _conv(g, sd, "downsample", fc, 1024, (1, 1, 1), bias=True)
    _conv(g, sd, "global_cls", cfg.num_classes, D + (0 if cfg.no_context else 1024), (1, 1, 1),
          bias=True, std=cls_std)
    _conv(g, sd, "local_conv.0.conv1", 1024, 832 + fc, (1, 1))
    _conv(g, sd, "local_conv.0.conv2", 256, 832 + fc, (1, 1))
    _conv(g, sd, "local_conv.0.conv3", 256, 256, (3, 3))
    _conv(g, sd, "local_conv.0.conv4", 1024, 256, (1, 1))
    for i in (1, 2):
        _conv(g, sd, "local_conv.%d.conv1" % i, 256, 1024, (1, 1))
        _conv(g, sd, "local_conv.%d.conv2" % i, 256, 256, (3, 3))
        _conv(g, sd, "local_conv.%d.conv3" % i, 1024, 256, (1, 1))
    _conv(g, sd, "downsample2", fc, 1024, (1, 1), bias=True)
    for name in ("local_reg", "neighbor_reg1", "neighbor_reg2"):
        sd[name + ".weight"] = torch.randn(4, D, generator=g) * reg_std
        sd[name + ".bias"] = torch.randn(4, generator=g) * 0.02
    return sd


def make_clips(B, T_in, H, W, seed=1234):
    """[B, T_in, 3, H, W] fp32 in [-1, 1] (the layout BaseNet.forward takes, networks.py:69-76)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T_in, 3, H, W, generator=g).clamp_(-1.0, 1.0)


def grid_anchors(scales=(4.0 / 3.0,), steps=(5.0 / 6.0,)):
    """Normalised grid anchors in the spirit of data/data_utils.py:19-45: for each (scale, step),
    boxes of side 1/scale... laid on a regular grid; here one scale -> 3x3 = 9 boxes."""
    out = []
    for s, st in zip(scales, steps):
        side = 1.0 / s
        n = 3
        for iy in range(n):
            for ix in range(n):
                cx = 0.5 + (ix - 1) * (1.0 - side) / 2.0 * st * 1.2
                cy = 0.5 + (iy - 1) * (1.0 - side) / 2.0 * st * 1.2
                out.append([max(cx - side / 2, 0.0), max(cy - side / 2, 0.0),
                            min(cx + side / 2, 1.0), min(cy + side / 2, 1.0)])
    return np.asarray(out, dtype=np.float32)


def make_proposals(B, N, T, W, H):
    """list (len B) of [N, T, 4] fp32 tubes in input pixels: 9 grid anchors + centre + full frame,
    cycled / truncated to N, replicated over T frames (the reference tiles anchors over T,
    data/ava.py:343-345)."""
    base = np.concatenate([grid_anchors(),
                           np.array([[0.25, 0.25, 0.75, 0.75], [0.0, 0.0, 1.0, 1.0]], np.float32)], 0)
    idx = np.arange(N) % base.shape[0]
    boxes = base[idx] * np.array([W, H, W, H], np.float32)
    jit = (np.arange(N) // base.shape[0]).astype(np.float32)[:, None] * 3.0  # distinct if N > 11
    boxes = boxes + jit * np.array([1, 1, -1, -1], np.float32)
    tubes = np.tile(boxes[:, None, :], (1, T, 1)).astype(np.float32)
    return [tubes.copy() for _ in range(B)]


def make_c3_rois(n_tubes=10000, n_clips=8, Tp=8, W=224, seed=0):
    """BASELINE config 3: random tubes -> flat ROI rows [n_tubes*Tp, 5] (frame index b*Tp+t first)
    and the per-tube boxes [n_tubes, 4] used for the NMS microbench."""
    rs = np.random.RandomState(seed)
    x1 = rs.uniform(0, 0.67 * W, n_tubes)
    y1 = rs.uniform(0, 0.67 * W, n_tubes)
    w = rs.uniform(0.09 * W, 0.54 * W, n_tubes)
    h = rs.uniform(0.09 * W, 0.54 * W, n_tubes)
    boxes = np.stack([x1, y1, np.minimum(x1 + w, W - 1), np.minimum(y1 + h, W - 1)], 1).astype(np.float32)
    clip = rs.randint(0, n_clips, n_tubes)
    frame = (clip[:, None] * Tp + np.arange(Tp)[None, :]).astype(np.float32)
    rois = np.concatenate([frame[:, :, None], np.tile(boxes[:, None, :], (1, Tp, 1))], 2).reshape(-1, 5)
    scores = rs.rand(n_tubes).astype(np.float32)
    return rois.astype(np.float32), boxes, scores


LOSS_CASES = {"c1": (3, 1, 5, "mixed"), "c3": (3, 3, 4, "mixed"), "nomask": (3, 1, 3, "zero")}


def make_loss_case(name, num_classes):
    """Seeded inputs of the training-time head call `TwoBranchNet.forward(feat, None, tubes=..., targets=...)`
    (two_branch.py:205-341): (T, feat [n,T',832,7,7], tubes [n,T',5], targets [n,3,6+cls]).  Shared by
    tests/golden/make_golden.py (reference run) and the oracle test."""
    T_, chunks, n, masks = LOSS_CASES[name]
    g = torch.Generator().manual_seed(77 + n)
    Tl = T_ * chunks
    feat = torch.randn(n, Tl, 832, 7, 7, generator=g) * 0.5
    x1 = torch.rand(n, Tl, generator=g) * 50
    y1 = torch.rand(n, Tl, generator=g) * 50
    w = 20 + torch.rand(n, Tl, generator=g) * 40
    h = 20 + torch.rand(n, Tl, generator=g) * 40
    tubes = torch.stack([torch.zeros(n, Tl), x1, y1, x1 + w, y1 + h], dim=2)
    tg = torch.zeros(n, 3, 6 + num_classes)
    gx1 = torch.rand(n, 3, generator=g) * 50
    gy1 = torch.rand(n, 3, generator=g) * 50
    tg[:, :, 0] = gx1
    tg[:, :, 1] = gy1
    tg[:, :, 2] = gx1 + 15 + torch.rand(n, 3, generator=g) * 45
    tg[:, :, 3] = gy1 + 15 + torch.rand(n, 3, generator=g) * 45
    if masks == "mixed":
        tg[:, :, 4] = (torch.rand(n, 3, generator=g) > 0.3).float()
        tg[:, :, 5] = (torch.rand(n, 3, generator=g) > 0.4).float()
        tg[0, :, 4] = 1.0
        tg[0, :, 5] = 1.0
    tg[:, :, 6:] = (torch.rand(n, 3, num_classes, generator=g) > 0.9).float()
    return T_, chunks, feat, tubes, tg
