"""TEST INFRASTRUCTURE ONLY -- functional torch-CPU restatement of the STEP model path.

The reference's model code *is* PyTorch modules (models/i3dpt.py, models/networks.py,
models/two_branch.py, utils/utils.py::inference); the arithmetic lives in PyTorch's CPU
backend (torch 2.11.0+cu128, oneDNN), which is the same library on this image and on the GPU box.
This file restates the reference's composition of those ops as plain functions over a state_dict
(reference key names), so that it can travel to the GPU box where /root/reference does not exist.
It is pinned against the reference modules themselves in tests/test_oracle_vs_reference.py
(build container only) and through tests/golden/*.npz fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops as _ops
from . import tubes as _tubes

BN_EPS = 1e-5  # torch.nn.BatchNorm3d default, i3dpt.py:101

# (in, [b0, b1a, b1b, b2a, b2b, b3]) -- i3dpt.py:213-231
MIXED = {
    "3b": (192, [64, 96, 128, 16, 32, 32]), "3c": (256, [128, 128, 192, 32, 96, 64]),
    "4b": (480, [192, 96, 208, 16, 48, 64]), "4c": (512, [160, 112, 224, 24, 64, 64]),
    "4d": (512, [128, 128, 256, 24, 64, 64]), "4e": (512, [112, 144, 288, 32, 64, 64]),
    "4f": (528, [256, 160, 320, 32, 128, 128]),
    "5b": (832, [256, 160, 320, 32, 128, 128]), "5c": (832, [384, 192, 384, 48, 128, 128]),
}


def same_pad(k, s):
    """i3dpt.py:14-31: pad_along = max(k - s, 0); low = pad//2; high = pad - low (per dim)."""
    pad = max(k - s, 0)
    return pad // 2, pad - pad // 2


def unit3d(x, sd, p, stride=(1, 1, 1), relu=True):
    """i3dpt.py:43-111: [ConstantPad3d 0] -> Conv3d(bias=False) -> BatchNorm3d(eval) -> ReLU."""
    w = sd[p + "conv3d.weight"]
    k = w.shape[2:]
    pads = [same_pad(k[i], stride[i]) for i in range(3)]  # (t, h, w)
    # F.pad order is (W_lo, W_hi, H_lo, H_hi, T_lo, T_hi)
    x = F.pad(x, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    y = F.conv3d(x, w, sd.get(p + "conv3d.bias"), stride=stride)
    if p + "batch3d.weight" in sd:
        y = F.batch_norm(y, sd[p + "batch3d.running_mean"], sd[p + "batch3d.running_var"],
                         sd[p + "batch3d.weight"], sd[p + "batch3d.bias"], False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def maxpool_tf(x, k, s):
    """i3dpt.py:114-126: zero ConstantPad3d (TF-SAME amounts) then MaxPool3d(ceil_mode=True)."""
    pads = [same_pad(k[i], s[i]) for i in range(3)]
    x = F.pad(x, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    return F.max_pool3d(x, k, s, ceil_mode=True)


def mixed(x, sd, p):
    """i3dpt.py:129-163."""
    b0 = unit3d(x, sd, p + "branch_0.")
    b1 = unit3d(unit3d(x, sd, p + "branch_1.0."), sd, p + "branch_1.1.")
    b2 = unit3d(unit3d(x, sd, p + "branch_2.0."), sd, p + "branch_2.1.")
    b3 = unit3d(maxpool_tf(x, (3, 3, 3), (1, 1, 1)), sd, p + "branch_3.1.")
    return torch.cat((b0, b1, b2, b3), 1)


def base_net(x, sd, prefix="base_model."):
    """networks.py:69-83 + 107-132.  x [N,T,C,H,W] -> [N,T/4,832,H/16,W/16]."""
    x = x.permute(0, 2, 1, 3, 4)
    p = prefix
    x = unit3d(x, sd, p + "0.", stride=(2, 2, 2))
    x = maxpool_tf(x, (1, 3, 3), (1, 2, 2))
    x = unit3d(x, sd, p + "2.")
    x = unit3d(x, sd, p + "3.")
    x = maxpool_tf(x, (1, 3, 3), (1, 2, 2))
    x = mixed(x, sd, p + "5.")
    x = mixed(x, sd, p + "6.")
    x = maxpool_tf(x, (3, 3, 3), (2, 2, 2))
    for i in range(8, 13):
        x = mixed(x, sd, p + "%d." % i)
    return x.permute(0, 2, 1, 3, 4)


def context_net(conv_feat, sd, prefix="i3d_conv_context.", global_mean=False):
    """two_branch.py:132-138.  The reference's AvgPool3d((1,13,13)) only accepts 400x400 inputs;
    global_mean=True is the resolution-free equivalent (identical at 400x400: 25 -> 13 -> 1)."""
    x = conv_feat.permute(0, 2, 1, 3, 4)
    x = maxpool_tf(x, (1, 3, 3), (1, 2, 2))
    x = mixed(x, sd, prefix + "1.")
    x = mixed(x, sd, prefix + "2.")
    if global_mean:
        return x.mean(dim=(3, 4), keepdim=True)
    return F.avg_pool3d(x, (1, 13, 13), (1, 1, 1))


def _bottleneck(x, sd, p, resample):
    """two_branch.py:60-111 (Conv2d, no BN, residual, ReLU)."""
    if resample:
        res = F.conv2d(x, sd[p + "conv1.weight"])
        o = F.relu(F.conv2d(x, sd[p + "conv2.weight"]))
        o = F.relu(F.conv2d(o, sd[p + "conv3.weight"], padding=1))
        o = F.conv2d(o, sd[p + "conv4.weight"])
    else:
        res = x
        o = F.relu(F.conv2d(x, sd[p + "conv1.weight"]))
        o = F.relu(F.conv2d(o, sd[p + "conv2.weight"], padding=1))
        o = F.conv2d(o, sd[p + "conv3.weight"])
    return F.relu(o + res)


def two_branch(global_feat, sd, T, context_feat=None, fc_dim=256, pool_size=7, cls_only=False, return_logits=False):
    """two_branch.py:205-274,337 in eval mode (dropout = identity), targets=None.
    global_feat [N,T',C,7,7] -> (global_prob [N,cls], local_loc [N,T',4], first_loc, last_loc)."""
    N, Tl, C, W, H = global_feat.shape
    chunks = int(Tl / T)
    chunk_idx = [j * T + int(T / 2) for j in range(chunks)]
    half_T = int(T / 2)
    g = global_feat.permute(0, 2, 1, 3, 4)
    g = mixed(g, sd, "i3d_conv.0.")
    g = mixed(g, sd, "i3d_conv.1.")
    gconv = F.conv3d(g, sd["downsample.weight"], sd["downsample.bias"])
    flat = gconv.permute(0, 2, 1, 3, 4).contiguous().view(N, Tl, -1, 1, 1).permute(0, 2, 1, 3, 4).contiguous()
    if context_feat is not None:
        flat = torch.cat([flat, context_feat], dim=1)
    cls = F.conv3d(flat, sd["global_cls.weight"], sd["global_cls.bias"]).squeeze(3).squeeze(3).mean(2)
    prob = torch.sigmoid(cls)
    if cls_only:
        z = torch.tensor([0.0])
        return (prob, z, z, z, cls) if return_logits else (prob, z, z, z)
    lf = torch.cat([global_feat.permute(0, 2, 1, 3, 4), gconv], dim=1)
    lf = lf.permute(0, 2, 1, 3, 4).contiguous().view(N * Tl, -1, W, H)
    lf = _bottleneck(lf, sd, "local_conv.0.", True)
    lf = _bottleneck(lf, sd, "local_conv.1.", False)
    lf = _bottleneck(lf, sd, "local_conv.2.", False)
    lf = F.conv2d(lf, sd["downsample2.weight"], sd["downsample2.bias"])
    lf = lf.reshape(lf.size(0), -1)
    local_loc = F.linear(lf, sd["local_reg.weight"], sd["local_reg.bias"]).view(N, Tl, -1)
    D = fc_dim * pool_size ** 2
    s0, s1 = chunk_idx[0] - half_T, chunk_idx[0] + half_T + 1
    e0, e1 = chunk_idx[-1] - half_T, chunk_idx[-1] + half_T + 1
    first = local_loc[:, s0:s1].contiguous().clone()
    last = local_loc[:, e0:e1].contiguous().clone()
    first = first + F.linear(lf.view(N, Tl, -1)[:, s0:s1].contiguous().view(-1, D),
                             sd["neighbor_reg1.weight"], sd["neighbor_reg1.bias"]).view(N, T, -1)
    last = last + F.linear(lf.view(N, Tl, -1)[:, e0:e1].contiguous().view(-1, D),
                           sd["neighbor_reg2.weight"], sd["neighbor_reg2.bias"]).view(N, T, -1)
    if return_logits:
        return prob, local_loc, first, last, cls
    return prob, local_loc, first, last


def encode_coef_t(gt, tubes):
    """tube_utils.py:143-163 on torch tensors."""
    def cs(b):
        w = b[:, 2] - b[:, 0] + 1.0
        h = b[:, 3] - b[:, 1] + 1.0
        return b[:, 0] + 0.5 * w, b[:, 1] + 0.5 * h, w, h
    gx, gy, gw, gh = cs(gt)
    x, y, w, h = cs(tubes)
    return torch.stack(((gx - x) / w, (gy - y) / h, torch.log(gw / w), torch.log(gh / h)), dim=1)


def two_branch_losses(cls_logits, local_loc, first_loc, last_loc, tubes, targets, T, cls_only=False):
    """two_branch.py:272-341 (training-time outputs; test infrastructure for the round-2 training path).
    cls_logits [N,cls] (pre-sigmoid), local_loc [N,T',4], first_loc/last_loc [N,T,4], tubes [N,T',5],
    targets [N,3,6+cls] = (first, centre, last) x (box 4 | cls mask | loc mask | labels).
    Returns (loss_global_cls, loss_local_loc, loss_neighbor_loc), each flattened like the reference's .view(-1)."""
    N, Tl = tubes.shape[0], tubes.shape[1]
    chunks = int(Tl / T)
    chunk_idx = [j * T + int(T / 2) for j in range(chunks)]
    half_T = int(T / 2)
    l_cls = torch.tensor(0.0)
    l_loc = torch.tensor(0.0)
    l_nb = torch.tensor(0.0)
    center_t, first_t, last_t = targets[:, 1].contiguous(), targets[:, 0].contiguous(), targets[:, -1].contiguous()
    center_tubes = tubes[:, chunk_idx[int(chunks / 2)]].contiguous()
    first_tubes = tubes[:, chunk_idx[0]].contiguous()
    last_tubes = tubes[:, chunk_idx[-1]].contiguous()
    mask = center_t[:, 4].view(-1, 1)
    if mask.sum():
        l_cls = F.binary_cross_entropy_with_logits(cls_logits, center_t[:, 6:] * mask, reduction="none")
    if not cls_only:
        center_pred = local_loc[:, chunk_idx[int(chunks / 2)]].contiguous().view(N, -1)
        first_pred = first_loc[:, half_T].contiguous().view(N, -1)
        last_pred = last_loc[:, half_T].contiguous().view(N, -1)
        tgt = encode_coef_t(center_t[:, :4].clone(), center_tubes.view(-1, 5)[:, 1:])
        m = center_t[:, 5].view(-1, 1).repeat(1, 4)
        if m.sum():
            l = F.smooth_l1_loss(center_pred, tgt, reduction="none")
            l_loc = torch.sum(l * m) / torch.sum(m)
        ntgt = encode_coef_t(torch.cat([first_t[:, :4], last_t[:, :4]], dim=0),
                             torch.cat([first_tubes.view(-1, 5)[:, 1:], last_tubes.view(-1, 5)[:, 1:]], dim=0))
        nm = torch.cat([first_t[:, 5].view(-1, 1).repeat(1, 4), last_t[:, 5].view(-1, 1).repeat(1, 4)], dim=0)
        if nm.sum():
            l = F.smooth_l1_loss(torch.cat([first_pred, last_pred], dim=0), ntgt, reduction="none")
            l_nb = torch.sum(l * nm) / torch.sum(nm)
    return l_cls.view(-1), l_loc.view(-1), l_nb.view(-1)


def roi_net(conv_feat, flat_tubes, pool_mode="align", pool_size=7, use_ref=True):
    """networks.py:34-47: ROIAlign((7,7), 1/16, 0) / ROIPool((7,7), 1/16) over [N*T', C, H, W]."""
    _, _, C, H, W = conv_feat.shape
    feat = conv_feat.reshape(-1, C, H, W).contiguous()
    rois = flat_tubes.reshape(-1, 5).contiguous()
    if pool_mode == "align":
        ref = _ops.ref_C() if use_ref else None
        if ref is not None:  # the reference's own compiled kernel when available
            return ref.roi_align_forward(feat, rois, 1.0 / 16.0, pool_size, pool_size, 0)
        return torch.from_numpy(_ops.roi_align_fwd(feat.numpy(), rois.numpy(), 1.0 / 16.0, pool_size, pool_size, 0))
    out, _ = _ops.roi_pool_fwd(feat.numpy(), rois.numpy(), 1.0 / 16.0, pool_size, pool_size)
    return torch.from_numpy(out)


def decode_coef_t(anchors, deltas):
    """tube_utils.py:165-189 on torch tensors (torch.exp, as the reference)."""
    w = anchors[:, 2] - anchors[:, 0] + 1.0
    h = anchors[:, 3] - anchors[:, 1] + 1.0
    x = anchors[:, 0] + 0.5 * w
    y = anchors[:, 1] + 0.5 * h
    px = w * deltas[:, 0] + x
    py = h * deltas[:, 1] + y
    pw = w * torch.exp(deltas[:, 2])
    ph = h * torch.exp(deltas[:, 3])
    out = deltas.clone()
    out[:, 0] = px - 0.5 * pw
    out[:, 1] = py - 0.5 * ph
    out[:, 2] = px + 0.5 * pw - 1
    out[:, 3] = py + 0.5 * ph - 1
    return out


def inference(args, conv_feat, context_feat, head_sds, exec_iter, tubes, pool_mode="align"):
    """utils/utils.py:15-131.  head_sds[i] = state_dict of det_net{i}.  Returns (history, trajectory)
    with the same dict keys as the reference."""
    flat, nums = _tubes.flatten_tubes(tubes, batch_idx=True)
    flat = torch.from_numpy(flat)
    history, trajectory = [], []
    for i in range(1, exec_iter + 1):
        chunks = args.NUM_CHUNKS[i]
        T_start = int((args.NUM_CHUNKS[args.max_iter] - chunks) / 2) * args.T
        T_len = chunks * args.T
        chunk_idx = [j * args.T + int(args.T / 2) for j in range(chunks)]
        half_T = int(args.T / 2)
        pooled = roi_net(conv_feat[:, T_start:T_start + T_len].contiguous(), flat, pool_mode, args.pool_size)
        _, C, W, H = pooled.shape
        pooled = pooled.view(-1, T_len, C, W, H)
        ctx = None
        if not args.no_context:
            ctx = torch.zeros((pooled.size(0), context_feat.size(1), T_len, 1, 1))
            for p in range(pooled.size(0)):
                ctx[p] = context_feat[int(flat[p, 0, 0].item() / T_len), :, T_start:T_start + T_len]
        prob, loc, first, last = two_branch(pooled, head_sds[i - 1], args.T, ctx, args.fc_dim, args.pool_size)
        pred_prob = prob.view(-1, 1, args.num_classes).expand(-1, T_len, -1)
        pred_loc = decode_coef_t(flat.view(-1, 5)[:, 1:], loc.reshape(-1, 4)).view(loc.size())
        pf = pl = None
        if args.temporal_mode == "predict":
            s0, s1 = chunk_idx[0] - half_T, chunk_idx[0] + half_T + 1
            e0, e1 = chunk_idx[-1] - half_T, chunk_idx[-1] + half_T + 1
            pf = decode_coef_t(flat[:, s0:s1].contiguous().view(-1, 5)[:, 1:], first.reshape(-1, 4)).view(first.size())
            pl = decode_coef_t(flat[:, e0:e1].contiguous().view(-1, 5)[:, 1:], last.reshape(-1, 4)).view(last.size())
        history.append({"pred_prob": pred_prob, "pred_loc": pred_loc, "pred_first_loc": pf,
                        "pred_last_loc": pl, "tubes_nums": nums})
        cur, selected, count = [], [], 0
        for b in range(len(nums)):
            s = count
            count += nums[b]
            cp = pred_prob[s:s + nums[b]]
            ct = pred_loc[s:s + nums[b]]
            cls = torch.argmax(cp, dim=-1)
            if i < args.max_iter and args.NUM_CHUNKS[i + 1] == args.NUM_CHUNKS[i] + 2:
                if args.temporal_mode == "predict":
                    prop = torch.cat([pf[s:s + nums[b]], ct, pl[s:s + nums[b]]], dim=1).numpy()
                elif args.temporal_mode == "extrapolate":
                    prop = _tubes.extrapolate_tubes(ct.numpy(), args.T)
                else:
                    prop = ct.numpy()
                    mean = np.tile(np.mean(prop, axis=1, keepdims=True), (1, args.T, 1))
                    prop = np.concatenate((mean, prop, mean), axis=1)
            else:
                prop = ct.numpy()
            prop = _tubes.valid_tubes(prop, width=args.image_size[0], height=args.image_size[1])
            cur.append((prop, cls))
            selected.append(prop)
        trajectory.append(cur)
        flat, nums = _tubes.flatten_tubes(selected, batch_idx=True)
        flat = torch.from_numpy(flat)
    return history, trajectory
