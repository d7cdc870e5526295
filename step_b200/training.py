"""First pieces of the training step on the device (the reference: train.py:286-348).

What exists: the head's three losses and the gradient of the training objective with respect to the head outputs
(`head_losses`), the backward of the head's small-N linears (`linear_backward`), a deterministic channels-last ROIAlign
backward (`roi_align_backward_nhwc`) and the tensor-core weight gradient of 1x1 convolutions (`conv1x1_wgrad`).
What does not exist yet: dgrad / wgrad of the k > 1 convolutions, max-pool backward, the optimizer step and the gradient
all-reduce -- `BaseNet` / `TwoBranchNet` therefore still run without autograd (their outputs carry no grad_fn).
"""
import torch

from . import _lib as L


def head_losses(logits, local_loc, first_loc, last_loc, tubes, targets, T, lambda_reg=5.0, lambda_neighbor=1.0,
                want_grads=False):
    """models/two_branch.py:276-333 on the device.  logits [N,cls] (pre-sigmoid), local_loc [N,T',4],
    first_loc / last_loc [N,Tc,4], tubes [N,T',5], targets [N,3,6+cls].
    Returns (loss_global_cls, loss_local_loc, loss_neighbor_loc) shaped like the reference's `.view(-1)` outputs
    (loss_global_cls is the element-wise BCE [N*cls], or the scalar 0 when no sample is positive) and, with
    want_grads, a dict with the gradients of  mean(loss_cls) + lambda_reg * loss_loc + lambda_neighbor * loss_nb
    (train.py:335-336; scripts/train_step.sh:43-44) w.r.t. logits / local_loc / first_loc / last_loc."""
    dev = L.same_device(logits, local_loc, first_loc, last_loc, tubes, targets)
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    logits, local_loc, first_loc, last_loc, tubes, targets = map(f32, (logits, local_loc, first_loc, last_loc, tubes, targets))
    N, cls = logits.shape
    T_len, Tc = local_loc.shape[1], first_loc.shape[1]
    if targets.shape != (N, 3, 6 + cls) or tubes.shape != (N, T_len, 5):
        raise RuntimeError("head_losses: targets %s / tubes %s do not match N=%d, T'=%d, classes=%d"
                           % (tuple(targets.shape), tuple(tubes.shape), N, T_len, cls))
    with torch.cuda.device(dev):
        loss_cls = torch.empty((N * cls,), dtype=torch.float32, device=dev)
        loss_loc = torch.empty((1,), dtype=torch.float32, device=dev)
        loss_nb = torch.empty((1,), dtype=torch.float32, device=dev)
        flags = torch.empty((3,), dtype=torch.int32, device=dev)
        scratch = torch.empty((N * 12,), dtype=torch.float32, device=dev)
        g = None
        if want_grads:
            g = {"logits": torch.empty_like(logits), "local_loc": torch.empty_like(local_loc),
                 "first_loc": torch.empty_like(first_loc), "last_loc": torch.empty_like(last_loc)}
        L.check(L.lib().step_head_losses_f32(L.ptr(logits), L.ptr(local_loc), L.ptr(first_loc), L.ptr(last_loc), L.ptr(tubes),
                                             L.ptr(targets), N, cls, T_len, int(T), Tc, float(lambda_reg), float(lambda_neighbor),
                                             L.ptr(loss_cls), L.ptr(loss_loc), L.ptr(loss_nb), L.ptr(flags),
                                             L.ptr(g["logits"]) if g else None, L.ptr(g["local_loc"]) if g else None,
                                             L.ptr(g["first_loc"]) if g else None, L.ptr(g["last_loc"]) if g else None,
                                             L.ptr(scratch), L.stream()))
        # `if mask.sum():` in the reference is the same host read-back (two_branch.py:293)
        if int(flags[0].item()) == 0:
            loss_cls = torch.zeros((1,), dtype=torch.float32, device=dev)
    return (loss_cls, loss_loc, loss_nb) if not want_grads else (loss_cls, loss_loc, loss_nb, g)


def linear_backward(x, w, dy, need_dx=True, need_dw=True, dx_out=None, accumulate_dx=False):
    """Backward of y = x W^T + b (nn.Linear, or the 1x1x1 `global_cls` on flattened features) for the head's small-N
    layers: x [M,K] fp16|fp32, w [Nn,K] fp32, dy [M,Nn] fp32 -> (dx [M,K] fp32 | None, dw [Nn,K] | None, db [Nn] | None)."""
    dev = L.same_device(x, w, dy)
    M, K = x.shape
    Nn = dy.shape[1]
    dy = dy.detach().float().contiguous()
    w32 = w.detach().float().contiguous() if w is not None else None
    with torch.cuda.device(dev):
        dx = None
        if need_dx:
            dx = dx_out if dx_out is not None else torch.empty((M, K), dtype=torch.float32, device=dev)
        dw = torch.empty((Nn, K), dtype=torch.float32, device=dev) if need_dw else None
        db = torch.empty((Nn,), dtype=torch.float32, device=dev) if need_dw else None
        xs = x.detach()
        if xs.stride(1) != 1:
            xs = xs.contiguous()
        L.check(L.lib().step_linear_small_n_bwd(L.ptr(xs), L.dt(xs), M, K, xs.stride(0), L.ptr(w32), L.ptr(dy), Nn, L.ptr(dx),
                                                1 if accumulate_dx else 0, L.ptr(dw), L.ptr(db), L.stream()))
    return dx, dw, db


def roi_align_backward_nhwc(grad_out, rois, spatial_scale, K, H, W, sampling_ratio=0):
    """grad_out [R,ph,pw,C] (channels-last, fp16|fp32) -> grad_in [K,H,W,C] fp32.  Deterministic: no atomics."""
    dev = L.same_device(grad_out, rois)
    R, ph, pw, C = grad_out.shape
    go = grad_out.detach().contiguous()
    r = rois.detach().float().contiguous()
    with torch.cuda.device(dev):
        gin = torch.empty((K, H, W, C), dtype=torch.float32, device=dev)
        L.check(L.lib().step_roi_align_bwd_nhwc(L.ptr(go), L.dt(go), C, L.ptr(r), R, float(spatial_scale), ph, pw, K, H, W, C,
                                                int(sampling_ratio), L.ptr(gin), C, L.stream()))
    return gin


def conv1x1_wgrad(dz, x, scale=1.0, out=None, accumulate=False):
    """dz [M,Cout] fp16, x [M,Cin] fp16 (row-major, channel strides allowed) -> dW [Cout,Cin] fp32 = scale * dz^T x."""
    dev = L.same_device(dz, x)
    if dz.dtype != torch.float16 or x.dtype != torch.float16:
        raise RuntimeError("conv1x1_wgrad: fp16 operands")
    M, Cout = dz.shape
    Cin = x.shape[1]
    with torch.cuda.device(dev):
        dw = out if out is not None else torch.empty((Cout, Cin), dtype=torch.float32, device=dev)
        nbytes = L.lib().step_conv1x1_wgrad_workspace_bytes(M, Cout, Cin)
        ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=dev)
        L.check(L.lib().step_conv1x1_wgrad_f16(L.ptr(dz), dz.stride(0), L.ptr(x), x.stride(0), M, Cout, Cin, float(scale), L.ptr(dw),
                                               dw.stride(0), 1 if accumulate else 0, L.ptr(ws), nbytes, L.stream()))
    return dw


# ---- backward of the tcgen05 / pool tape ------------------------------------------------------------------------------
class GradStore:
    """fp16 gradient buffers, one per activation buffer of the forward (same shape, zero-initialised on first use).
    Every consumer ACCUMULATES into the slice it read, so fan-out (Inception branches, residuals, the shared concat
    buffer of two_branch.py:256) needs no special casing."""

    def __init__(self):
        self.bufs = {}

    def of(self, act):
        """Act view of the gradient of `act` (same channel slice of the gradient buffer)."""
        from .engine import Act
        key = act.buf.data_ptr()
        g = self.bufs.get(key)
        if g is None:
            g = torch.zeros_like(act.buf)
            self.bufs[key] = g
        # `act` may be another view of the same memory (Act.frames(): [N*T, 1, H, W, ld] over [N, T, H, W, ld])
        return Act(g.view(act.buf.shape), act.C, act.coff)


def _dgrad_weights(entry):
    """[Cout, taps, cin_pad] forward weights -> [Cin, taps, Cout] for the input gradient: dx = conv(dz, flip(w)^T)."""
    c = entry.get("_wT")
    if c is None:
        cin = entry["x"].C
        c = entry["w"][:, :, :cin].flip(1).permute(2, 1, 0).contiguous()
        entry["_wT"] = c
    return c


TIMING = None     # set to {} to collect per-phase device times of tape_backward (tools/train_bench.py)


class _Phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if TIMING is not None:
            self.e1.record()
            TIMING.setdefault(self.name, []).append((self.e0, self.e1))


def timing_summary():
    torch.cuda.synchronize()
    return {k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in (TIMING or {}).items()}


def tape_backward(tape, grads, loss_scale=1.0, need_input_grad=None):
    """Reverse pass over the conv / max-pool tape recorded by engine.TAPE (fp16 path).  `grads` (GradStore) must already
    hold d(loss * loss_scale)/d(output) of the last layers.  Returns {parameter tensor: fp32 gradient in the parameter's
    own layout} (BatchNorm is frozen, i3dpt / networks.py:136-142: only conv weights and biases train).
    The input gradient of every layer is produced by the SAME tcgen05 convolution kernels run on the transposed, flipped
    filter (stride-1 convolutions: dx = conv(dz, flip(w)^T)), accumulated through their residual input."""
    from . import engine as E
    from .engine import Act
    out = {}
    lib = L.lib()
    inv = 1.0 / float(loss_scale)

    def add(param, g):
        out[param] = out[param] + g if param in out else g

    for e in reversed(tape):
        if e["kind"] == "pool":
            x, y = e["x"], e["out"]
            gy, gx = grads.of(y), grads.of(x)
            k, s, pl, ph = e["k"], e["stride"], e["pad_lo"], e["pad_hi"]
            ws = torch.empty((y.N * y.T * y.H * y.W * y.C,), dtype=torch.uint8, device=x.device)
            with _Phase("pool_bwd"):
              L.check(lib.step_maxpool3d_bwd_f16(L.c_void_p(x.data_ptr()), x.ld, L.c_void_p(gy.data_ptr()), gy.ld, x.N, x.T, x.H, x.W,
                                               x.C, k[0], k[1], k[2], s[0], s[1], s[2], pl[0], pl[1], pl[2], ph[0], ph[1], ph[2],
                                               y.T, y.H, y.W, L.c_void_p(gx.data_ptr()), gx.ld, L.ptr(ws), L.stream()))
            continue
        x, w, k = e["x"], e["w"], e["k"]
        if e["stride"] != (1, 1, 1):
            raise NotImplementedError("tape_backward: strided convolution (the trunk's stem) is not built")
        outs = [e["out"]] + e["extra_outs"]
        M = x.N * x.T * x.H * x.W
        n_total = sum(o.C for o in outs)
        dz = torch.empty((x.N, x.T, x.H, x.W, n_total), dtype=torch.float16, device=x.device)
        col = 0
        ph_act = _Phase("act_bwd"); ph_act.__enter__()
        for o in outs:
            gy = grads.of(o)
            sc = e["scale"][col:col + o.C] if e["scale"] is not None else None
            res = e["residual"]
            gres = grads.of(res) if res is not None else None
            L.check(lib.step_act_bwd_f16(L.c_void_p(gy.data_ptr()), gy.ld, L.c_void_p(o.data_ptr()), o.ld, L.ptr(sc), 1 if e["relu"] else 0,
                                         M, o.C, L.c_void_p(dz.data_ptr() + 2 * col), n_total,
                                         L.c_void_p(gres.data_ptr()) if gres is not None else None, gres.ld if gres is not None else 0,
                                         L.stream()))
            col += o.C
        ph_act.__exit__()
        # ---- weight (and bias) gradients
        taps = k[0] * k[1] * k[2]
        dw = torch.empty((n_total, taps, x.C), dtype=torch.float32, device=x.device)
        nbytes = lib.step_conv_wgrad_workspace_bytes(M, n_total, x.C, taps)
        ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=x.device)
        pl = e["pad_lo"]
        with _Phase("wgrad_k%d" % taps):
          L.check(lib.step_conv_wgrad_f16(L.ptr(dz), n_total, L.c_void_p(x.data_ptr()), x.ld, x.N, x.T, x.H, x.W, n_total, x.C, k[0], k[1],
                                        k[2], pl[0], pl[1], pl[2], inv, L.ptr(dw), x.C, 0, L.ptr(ws), nbytes, L.stream()))
        if isinstance(e["tag"], tuple) and e["tag"][0] == "s2d":
            # the stride-2 7x7x7 stem runs as a 4x4x4 filter over the space-to-depth clip (engine.pack_stem_s2d):
            # tap q and sub-position r hold filter position k = 2 q + r (k = 7 is padding) -- undo that packing
            unit = e["tag"][1]
            cin = unit.conv3d.in_channels
            g8 = dw[:, :, :8 * cin].reshape(n_total, 4, 4, 4, 2, 2, 2, cin).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(n_total, cin, 8, 8, 8)
            if unit.conv3d.weight.requires_grad:
                add(unit.conv3d.weight, g8[:, :, :7, :7, :7].contiguous())
            continue                                   # the clip itself needs no gradient
        tags = e["tag"] if isinstance(e["tag"], (list, tuple)) else [e["tag"]]
        row = 0
        for tg, o in zip(tags, outs):
            if tg is None:
                row += o.C
                continue
            conv = getattr(tg, "conv3d", tg)          # Unit3Dpy holds its nn.Conv3d; nn.Conv2d / nn.Conv3d containers are themselves
            g = dw[row:row + o.C].view((o.C,) + tuple(conv.weight.shape[2:]) + (x.C,))
            g = g.permute(0, g.dim() - 1, *range(1, g.dim() - 1)).contiguous()        # [Cout, Cin, *k]
            if conv.weight.requires_grad:
                add(conv.weight, g)
            if conv.bias is not None and conv.bias.requires_grad:
                db = torch.empty((o.C,), dtype=torch.float32, device=x.device)
                wsb = torch.empty((64 * o.C,), dtype=torch.float32, device=x.device)
                L.check(lib.step_colsum_f16(L.c_void_p(dz.data_ptr() + 2 * row), n_total, M, o.C, inv, L.ptr(db), L.ptr(wsb), L.stream()))
                add(conv.bias, db)
            row += o.C
        # ---- input gradient: the forward kernel on the transposed, flipped filter, accumulated into grad(x)
        if need_input_grad is None or need_input_grad(e):
            gx = grads.of(x)
            wT = _dgrad_weights(e)
            pad = tuple(kk - 1 - p for kk, p in zip(k, pl))
            saved, E.TAPE = E.TAPE, None
            try:
                with _Phase("dgrad"):
                    E.conv(Act(dz), wT, None, None, gx, k, (1, 1, 1), pad, relu=False, residual=gx, out_dims=(x.T, x.H, x.W))
            finally:
                E.TAPE = saved
    return out


def head_forward_backward(net, global_feat, tubes, targets, context_feat=None, lambda_reg=5.0, lambda_neighbor=1.0,
                          loss_scale=1024.0, cat=None, objective_scale=1.0):
    """One training-time evaluation of a TwoBranchNet on the device (train.py:323-347 for one refinement step): forward
    with targets, the three losses, and the gradient of  mean(loss_cls) + lambda_reg * loss_loc + lambda_neighbor * loss_nb
    with respect to every trainable parameter of the head and to the pooled ROI features.  fp16 activations / activation
    gradients with a static loss scale (apex-style), fp32 weight gradients.  Dropout is the identity (eval mode), like the
    reference's gradient goldens.
    Returns dict(prob, loc, first, last, losses=(cls, loc, nb), loss, grads={param: grad}, feat_grad=[N,T',832,7,7] fp32)."""
    from . import engine as E
    from .engine import Act
    from .networks import to_act
    if context_feat is not None:
        raise NotImplementedError("head_forward_backward: the context branch's backward is not built")
    if E.dtype_code(net.fp16) != L.F16:
        raise RuntimeError("head_forward_backward runs on the fp16 path (cfg.fp16=True)")
    fc, ps = net.fc_dim, net.pool_size
    D = fc * ps * ps
    if cat is None:
        dev = L.same_device(global_feat, tubes, targets)
        N, Tl, C, Wd, Hd = global_feat.shape
    else:   # ROI features already pooled into the [ROI | downsample] concat buffer (ROINet.pool_into)
        dev = L.same_device(cat.buf, tubes, targets)
        N, Tl, Wd, Hd, C = cat.N, cat.T, cat.H, cat.W, cat.ld - fc
    with torch.cuda.device(dev), torch.no_grad():
        if cat is None:
            cat = Act.empty(N, Tl, Wd, Hd, C + fc, L.F16, dev)
            src = to_act(global_feat, L.F16)
            cat.buf[..., :C].copy_(src.buf[..., src.coff:src.coff + C])
        tape, keep = [], {}
        saved_tape, saved_bs = E.TAPE, E.BRANCH_STREAMS
        E.TAPE, E.BRANCH_STREAMS = tape, False          # one stream: the tape order is the execution order
        try:
            prob, loc, first, last, logits = net.forward_act(cat, None, None, want_logits=True, keep=keep)
        finally:
            E.TAPE, E.BRANCH_STREAMS = saved_tape, saved_bs
        lc, ll, ln, g = head_losses(logits, loc, first, last, tubes, targets, net.T, lambda_reg, lambda_neighbor, want_grads=True)
        if objective_scale != 1.0:
            for k_ in g:
                g[k_].mul_(objective_scale)
        grads = GradStore()
        out = {}
        hw = net._head_weights()
        unperm = lambda w: w.view(-1, ps * ps, fc).permute(0, 2, 1).reshape(w.shape[0], -1)   # (p*fc + c) -> (c*49 + p)
        # ---- classifier: logits = mean_t(gconv) . W^T + b   (two_branch.py:246-249)
        dxbar, dw, db = linear_backward(keep["xbar"], hw["cls_w"], g["logits"])
        out[net.global_cls.weight] = unperm(dw).reshape(net.global_cls.weight.shape)
        out[net.global_cls.bias] = db
        gcat = grads.of(cat)
        L.check(L.lib().step_mean_mid_bwd(L.ptr(dxbar), N, Tl, ps * ps, fc, float(loss_scale),
                                          L.c_void_p(gcat.data_ptr() + 2 * C), gcat.ld, L.stream()))
        # ---- regressors (two_branch.py:261-270): local_reg on every frame, neighbor_reg1 / 2 on the first / last chunk
        lf2 = keep["local_feat2"]
        lf2v = lf2.buf.view(N, Tl, D)
        s0, s1, e0, e1 = keep["slices"]
        dlf2 = torch.zeros((N, Tl, D), dtype=torch.float32, device=dev)
        dx, dw, db = linear_backward(lf2v.reshape(N * Tl, D), hw["local_reg_w32"], g["local_loc"].reshape(N * Tl, 4), dx_out=dlf2.view(N * Tl, D))
        out[net.local_reg.weight], out[net.local_reg.bias] = unperm(dw), db
        for mod, nm, (a, b), gk in ((net.neighbor_reg1, "neighbor_reg1", (s0, s1), "first_loc"), (net.neighbor_reg2, "neighbor_reg2", (e0, e1), "last_loc")):
            xs = lf2v[:, a:b].reshape(-1, D).contiguous()
            dx, dw, db = linear_backward(xs, hw[nm + "_w32"], g[gk].reshape(-1, 4))
            dlf2[:, a:b] += dx.view(N, b - a, D)                      # disjoint frame ranges of one buffer (host-side glue)
            out[mod.weight], out[mod.bias] = unperm(dw), db
        glf2 = grads.of(lf2)
        L.check(L.lib().step_f32_accum_f16(L.ptr(dlf2), N * Tl * ps * ps, fc, float(loss_scale), L.c_void_p(glf2.data_ptr()), glf2.ld,
                                           L.stream()))
        # ---- every convolution and pool of the head, in reverse
        out.update(tape_backward(tape, grads, loss_scale))
        gcat = grads.of(cat)
        fg = gcat.buf[..., :C].float().mul_(1.0 / loss_scale).permute(0, 1, 4, 2, 3).contiguous() if global_feat is not None else None
    loss = lc.mean() + lambda_reg * ll.mean() + lambda_neighbor * ln.mean()
    return dict(prob=prob, loc=loc, first=first, last=last, losses=(lc, ll, ln), loss=loss, grads=out, feat_grad=fg,
                roi_grad=Act(gcat.buf, C, 0))


def trunk_forward_backward(base_net, clips, d_feat_fn, loss_scale=1024.0):
    """I3D trunk forward on the fp16 path with the tape on, then backward from d(loss)/d(conv_feat).
    d_feat_fn(feat_act) -> fp32 tensor [N,T',H',W',832] (channels-last) with the gradient of the loss w.r.t. the trunk
    output (e.g. the sum of the ROIAlign backward results of the refinement steps).
    Returns (feat Act, {conv weight: fp32 gradient}).  BatchNorm stays frozen (networks.py:85-99,136-142)."""
    from . import engine as E
    dev = clips.device
    with torch.cuda.device(dev), torch.no_grad():
        tape = []
        saved_tape, saved_bs = E.TAPE, E.BRANCH_STREAMS
        E.TAPE, E.BRANCH_STREAMS = tape, False
        try:
            feat = base_net.forward_act(clips)
        finally:
            E.TAPE, E.BRANCH_STREAMS = saved_tape, saved_bs
        grads = GradStore()
        gfeat = grads.of(feat)
        d = d_feat_fn(feat).to(torch.float32).contiguous()
        M = feat.N * feat.T * feat.H * feat.W
        L.check(L.lib().step_f32_accum_f16(L.ptr(d), M, feat.C, float(loss_scale), L.c_void_p(gfeat.data_ptr()), gfeat.ld, L.stream()))
        out = tape_backward(tape, grads, loss_scale)
    return feat, out


def sgd_step(params_and_grads, lr, momentum=0.9, weight_decay=0.0, state=None, world_size=1):
    """optim.SGD(momentum, weight_decay) (train.py:124) on the fp32 master parameters, after an optional gradient
    all-reduce over the clip-parallel ranks (NCCL; one flat bucket).  state: dict param -> momentum buffer.
    Host-side glue over torch.distributed + elementwise updates; the parameters change in place (their packed fp16
    copies are rebuilt by the modules' version-keyed caches on the next forward)."""
    state = {} if state is None else state
    items = [(p, g) for p, g in params_and_grads.items() if p.requires_grad]
    if world_size > 1 and items:
        import torch.distributed as dist
        flat = torch.cat([g.reshape(-1) for _, g in items])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world_size)
        off = 0
        for _, g in items:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    with torch.no_grad():
        for p, g in items:
            g = g.to(p.device, p.dtype)
            if weight_decay:
                g = g.add(p, alpha=weight_decay)
            if momentum:
                buf = state.get(p)
                if buf is None:
                    buf = state[p] = g.clone()          # torch.optim.SGD: first step initialises the buffer with the gradient
                else:
                    buf.mul_(momentum).add_(g)
                g = buf
            p.add_(g, alpha=-lr)
    return state


def train_step(cfg, nets, clips, step_tubes, step_targets, lr=None, momentum=0.9, weight_decay=0.0, lambda_reg=5.0,
               lambda_neighbor=1.0, loss_scale=1024.0, sgd_state=None, world_size=1):
    """One optimisation step of train.py:286-348 on the device, for already selected training samples
    (`train_select`, utils/utils.py:135-423, is the host-side sampling of SURVEY.md section 8f rank 4 and is not built):
        conv_feat = base_net(clips)                                   train.py:263
        for each refinement step i: ROI-pool the step's flat tubes, run det_net[i-1] with targets,
            loss_back += loss_cls.mean() + lambda_reg * loss_loc.mean() + lambda_neighbor * loss_nb.mean()   train.py:323-336
        loss_back.backward(); optimizer.step()                         train.py:345-348
    step_tubes[i]: [R_i, T', 5] fp32 (frame index first, as flatten_tubes(batch_idx=True) builds them), step_targets[i]:
    [R_i, 3, 6 + classes].  Spatial mode only (every step pools the whole T' range).  Context branch off.
    Returns dict(loss, losses=[(cls, loc, nb)], grads={param: fp32 grad}); updates the parameters when lr is given."""
    from . import engine as E
    from .engine import Act
    base, roi_net = nets["base_net"], nets["roi_net"]
    dev = clips.device
    n_steps = len(step_tubes)
    results, roi_grads = [], []
    all_grads = {}

    def d_feat(feat):
        # heads first (they need conv_feat), then the sum of their ROIAlign backward results is the trunk's output gradient
        total = None
        for i in range(n_steps):
            head = nets["det_net%d" % i]
            flat = step_tubes[i].to(dev).float().contiguous()
            R, Tl = flat.shape[0], flat.shape[1]
            if Tl != feat.T:
                raise NotImplementedError("train_step: temporal chunking of the training step is not built (spatial mode)")
            cat = Act.empty(R, Tl, head.pool_size, head.pool_size, 832 + head.fc_dim, L.F16, dev)
            roi_net.pool_into(feat, flat, cat.frames().slice(0, 832), Tl, feat.T, 0)
            r = head_forward_backward(head, None, flat, step_targets[i].to(dev), lambda_reg=lambda_reg, lambda_neighbor=lambda_neighbor,
                                      loss_scale=loss_scale, cat=cat)
            results.append(r)
            all_grads.update(r["grads"])
            rg = r["roi_grad"]                                     # fp16, scaled by loss_scale, [R, T', 7, 7, ld] slice [0, 832)
            gin = roi_align_backward_nhwc_strided(rg, flat.view(-1, 5), 1.0 / 16.0, feat.N * feat.T, feat.H, feat.W)
            total = gin if total is None else total.add_(gin)
        return total.mul_(1.0 / loss_scale).view(feat.N, feat.T, feat.H, feat.W, feat.C)

    feat, tg = trunk_forward_backward(base, clips, d_feat, loss_scale)
    all_grads.update(tg)
    loss = sum(r["loss"] for r in results)
    if lr is not None:
        sgd_state = sgd_step(all_grads, lr, momentum, weight_decay, sgd_state, world_size)
    return dict(loss=loss, losses=[r["losses"] for r in results], grads=all_grads, sgd_state=sgd_state)


def roi_align_backward_nhwc_strided(grad_act, rois, spatial_scale, K, H, W, sampling_ratio=0):
    """ROIAlign backward from a channel slice of a wider fp16 buffer (Act [R, T', ph, pw, ld] slice of C channels)."""
    ph, pw, C = grad_act.H, grad_act.W, grad_act.C
    R = grad_act.N * grad_act.T
    dev = grad_act.device
    r = rois.detach().float().contiguous()
    gin = torch.empty((K, H, W, C), dtype=torch.float32, device=dev)
    L.check(L.lib().step_roi_align_bwd_nhwc(L.c_void_p(grad_act.data_ptr()), grad_act.code, grad_act.ld, L.ptr(r), R, float(spatial_scale),
                                            ph, pw, K, H, W, C, int(sampling_ratio), L.ptr(gin), C, L.stream()))
    return gin
