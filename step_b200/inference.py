"""`inference()` -- the progressive-refinement loop of the reference (utils/utils.py:15-131) with the
same signature and return structure, kept entirely on the device.

Per step the reference does: flatten_tubes (numpy) -> H2D -> ROI pool -> per-tube Python loop with one
`.item()` sync per tube for the context row -> head -> decode -> per-clip D2H -> numpy
extrapolate/concat -> per-box Python loop in valid_tubes -> flatten -> H2D.  Here a step is: one
ROIAlign launch writing straight into the head's concat buffer, the head, and ONE fused
`tube_update` launch (decode + extension + validation + re-flatten).  Nothing crosses PCIe until the
caller reads the history.
"""
import numpy as np
import torch

from . import _lib as L
from . import engine as E
from .engine import Act
from .networks import act_of, to_act
from .tube_utils import flatten_tubes, tube_update


def _ext_mode(args, i):
    if i < args.max_iter and args.NUM_CHUNKS[i + 1] == args.NUM_CHUNKS[i] + 2:
        return {"predict": L.EXT_PREDICT, "extrapolate": L.EXT_EXTRAPOLATE}.get(args.temporal_mode, L.EXT_MEAN)
    return L.EXT_NONE


def stage_tubes(tubes, device):
    """Host list of [n_b, T, 4] numpy tubes -> (flat [R,T,5] fp32, clip_of_tube [R] int32, tubes_nums) on device.
    flatten_tubes is list bookkeeping only (tube_utils.py:214-246); this is the one H2D of the loop."""
    flat_np, tubes_nums = flatten_tubes(tubes, batch_idx=True)
    flat = torch.from_numpy(np.ascontiguousarray(flat_np, dtype=np.float32)).to(device)
    clip_of_tube = torch.from_numpy(np.repeat(np.arange(len(tubes_nums)), tubes_nums).astype(np.int32)).to(device)
    return flat, clip_of_tube, tubes_nums


def inference_device(args, feat, ctx_all, nets, exec_iter, flat, clip_of_tube, tubes_nums):
    """Device-only body of the loop (no host<->device traffic, no synchronisation, CUDA-graph
    capturable).  feat: Act [B,T',H',W',832]; ctx_all: fp32 [B,T',1024] or None; flat [R,T,5].
    Returns (history, [(flat_next, prob)] per step)."""
    dev = feat.device
    code = feat.code
    B, T_total = feat.N, feat.T
    R = flat.shape[0]
    roi_net = nets['roi_net']
    history, steps = [], []
    width, height = float(args.image_size[0]), float(args.image_size[1])
    decode_nb = args.temporal_mode == "predict"
    with torch.cuda.device(dev):
        for i in range(1, exec_iter + 1):
            chunks = args.NUM_CHUNKS[i]
            T_start = int((args.NUM_CHUNKS[args.max_iter] - chunks) / 2) * args.T
            T_len = chunks * args.T
            if flat.shape[1] != T_len:
                raise RuntimeError("inference: tubes have %d frames but step %d pools %d" % (flat.shape[1], i, T_len))
            head = nets['det_net%d' % (i - 1)]
            ps = head.pool_size
            # ROI pooling straight into [ROI feat | downsample] concat buffer (utils.py:48, two_branch.py:256)
            cat = Act.empty(R, T_len, ps, ps, 832 + head.fc_dim, code, dev)
            roi_net.pool_into(feat, flat, cat.frames().slice(0, 832), T_len, T_total, T_start)
            ctx_mean = None
            if ctx_all is not None:
                sl = ctx_all[:, T_start:T_start + T_len].contiguous()
                ctx_mean = E.mean_mid(sl.data_ptr(), L.F32, B, T_len, 1, sl.shape[2], sl.shape[2], dev)
            # test.py:85-87 may have placed this head on another GPU (set_device): pool here, run the head there
            # (two_branch.py:225-229 moves the pooled feature the same way) and bring the four small results back
            hdev = torch.device(head.device) if getattr(head, "device", None) not in (None, "cpu") else dev
            if hdev.type == "cuda" and hdev.index is None:
                hdev = torch.device("cuda", torch.cuda.current_device())
            if hdev != dev:
                cat_h = Act(cat.buf.to(hdev))
                cm = ctx_mean.to(hdev) if ctx_mean is not None else None
                with torch.cuda.device(hdev):
                    outs = head.forward_act(cat_h, cm, clip_of_tube.to(hdev))
                prob, loc, first, last = (o.to(dev) for o in outs)
            else:
                prob, loc, first, last = head.forward_act(cat, ctx_mean, clip_of_tube)
            ext = _ext_mode(args, i)
            pred_loc, pf, pl, flat_next = tube_update(flat, loc, first if decode_nb else None,
                                                      last if decode_nb else None, clip_of_tube, args.T, decode_nb,
                                                      ext, width, height)
            pred_prob = prob.view(-1, 1, args.num_classes).expand(-1, T_len, -1)
            history.append({'pred_prob': pred_prob, 'pred_loc': pred_loc, 'pred_first_loc': pf, 'pred_last_loc': pl,
                            'tubes_nums': tubes_nums})
            steps.append((flat_next, prob))
            flat = flat_next
    return history, steps


def inference(args, conv_feat, context_feat, nets, exec_iter, tubes, want_trajectory=True):
    """Same contract as utils/utils.py:15-131.

    conv_feat: logical [B, T', 832, H', W'] tensor from BaseNet (a view of our channels-last buffer;
    any other CUDA tensor of that shape is converted once).  context_feat: [B, 1024, T', 1, 1] or None.
    nets: {'roi_net': ROINet, 'det_net%d': TwoBranchNet}.  tubes: list of [n_b, T, 4] numpy arrays.

    Returns (history, trajectory): history[i] = {'pred_prob' [R,T_len,cls] (expand view),
    'pred_loc' [R,T_len,4], 'pred_first_loc', 'pred_last_loc' ([R,T,4] or None), 'tubes_nums'};
    trajectory[i][b] = (proposals numpy [n_b,T_next,4], pred_class tensor) when want_trajectory.
    """
    L.need_cuda(conv_feat)
    dev = conv_feat.device
    code = E.dtype_code(nets['det_net0'].fp16)
    feat = act_of(conv_feat)
    if feat is None or feat.code != code:
        feat = to_act(conv_feat, code)
    flat, clip_of_tube, tubes_nums = stage_tubes(tubes, dev)
    ctx_all = None
    if not args.no_context:
        L.need_cuda(context_feat)
        ctx_all = context_feat.detach().float().reshape(feat.N, context_feat.shape[1], feat.T).permute(0, 2, 1).contiguous()
    history, steps = inference_device(args, feat, ctx_all, nets, exec_iter, flat, clip_of_tube, tubes_nums)

    trajectory = []
    if want_trajectory:  # one synchronisation at the very end instead of B per step
        for (flat_next, prob), h in zip(steps, history):
            props = flat_next[:, :, 1:].cpu().numpy()
            cls = torch.argmax(prob, dim=-1).cpu()
            T_len = h['pred_loc'].shape[1]     # utils.py:95: argmax over pred_prob, which has this step's T_len frames
            cur, s = [], 0
            for n in tubes_nums:
                cur.append((props[s:s + n], cls[s:s + n].view(-1, 1).expand(-1, T_len)))
                s += n
            trajectory.append(cur)
    return history, trajectory
