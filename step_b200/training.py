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
