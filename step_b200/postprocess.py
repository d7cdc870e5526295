"""Detection post-processing on the device: the per-clip x per-class loop of the reference drivers
(test.py:156-218, demo.py:121-198 -- B*60*max_iter tiny CPU NMS calls per batch) as ONE segmented launch.

For every (clip, class): keep tubes whose centre-frame score > conf_thresh, clamp their centre-frame box
with valid_tubes (the drivers call it with its 400x400 defaults, test.py:191), run greedy NMS
(cpu/nms_cpu.cpp semantics, bit-exact), normalise by (width, height), then take the top-k scores per clip.
"""
import numpy as np
import torch

from . import tube_utils
from .roi_layers import nms_segmented

_plan_cache = {}


def _plan(tubes_nums, num_classes, device):
    key = (tuple(tubes_nums), num_classes, str(device))
    p = _plan_cache.get(key)
    if p is None:
        rows, cls, offs, clip = [], [], [0], []
        start = 0
        for b, n in enumerate(tubes_nums):
            for c in range(num_classes):
                rows.append(np.arange(start, start + n))
                cls.append(np.full(n, c))
                clip.append(np.full(n, b))
                offs.append(offs[-1] + n)
            start += n
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
        p = tuple(torch.from_numpy(a).to(device) for a in
                  (cat(rows).astype(np.int64), cat(cls).astype(np.int64), np.asarray(offs, np.int32), cat(clip).astype(np.int64)))
        _plan_cache[key] = p
    return p


def detect(pred_prob, pred_loc, tubes_nums, conf_thresh, nms_thresh, width, height, topk=0, valid_size=(400, 400)):
    """pred_prob [R,T,cls] | [R,cls], pred_loc [R,T,4] (history[i] of `inference`), all on the device.
    Returns a dict of device tensors over the R*cls candidate rows (clip-major, class, tube):
    keep (bool), score, box (normalised x1,y1,x2,y2), tube (row of pred_*), cls, clip -- no synchronisation.
    With topk > 0, `keep` is further restricted to the k best kept scores of every clip (test.py:205-208)."""
    prob = pred_prob[:, pred_prob.shape[1] // 2] if pred_prob.dim() == 3 else pred_prob
    boxes = pred_loc[:, pred_loc.shape[1] // 2].contiguous().clone()
    tube_utils.valid_tubes(boxes.view(-1, 1, 4), width=valid_size[0], height=valid_size[1])
    rows, cls, offs, clip = _plan(tubes_nums, prob.shape[1], prob.device)
    score = prob.float()[rows, cls].contiguous()
    cand = boxes[rows].contiguous()
    strictly_greater = float(np.nextafter(np.float32(conf_thresh), np.float32(np.inf)))   # scores.gt(conf_thresh)
    keep = nms_segmented(cand, score, offs, nms_thresh, min_score=strictly_greater).bool()
    if topk and topk > 0:
        masked = torch.where(keep, score, torch.full_like(score, -1.0))
        ncls = prob.shape[1]
        start = 0
        for n in tubes_nums:  # candidate rows of a clip are contiguous: n * ncls of them (views, no sync)
            cnt = n * ncls
            if cnt > topk:
                seg = masked.narrow(0, start, cnt)
                thr = torch.topk(seg, topk).values[-1]
                keep.narrow(0, start, cnt).logical_and_(seg >= thr)
            start += cnt
    norm = torch.tensor([width, height, width, height], dtype=torch.float32, device=cand.device)
    return {"keep": keep, "score": score, "box": cand / norm, "tube": rows, "cls": cls, "clip": clip}


def to_lists(det, n_clips):
    """Materialise (one D2H) as per-clip lists of (box[4], cls, score), best score first."""
    k = det["keep"].cpu().numpy()
    sc, bx = det["score"].cpu().numpy()[k], det["box"].cpu().numpy()[k]
    cl, cp = det["cls"].cpu().numpy()[k], det["clip"].cpu().numpy()[k]
    out = []
    for b in range(n_clips):
        m = np.nonzero(cp == b)[0]
        m = m[np.argsort(-sc[m], kind="stable")]
        out.append([(bx[i], int(cl[i]), float(sc[i])) for i in m])
    return out
