"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the tube arithmetic on the STEP hot path.

Follows /root/reference/utils/tube_utils.py (line numbers per function).  All arrays fp32.
Pinned against the reference functions themselves in tests/test_oracle_vs_reference.py (build
container) and through tests/golden/tubes_*.npz everywhere else.
"""
import numpy as np

F = np.float32


def get_center_size(boxes):
    """tube_utils.py:127-141: w = x2-x1+1, h = y2-y1+1, x = x1+.5w, y = y1+.5h."""
    boxes = np.asarray(boxes, dtype=F)
    w = boxes[:, 2] - boxes[:, 0] + F(1.0)
    h = boxes[:, 3] - boxes[:, 1] + F(1.0)
    x = boxes[:, 0] + F(0.5) * w
    y = boxes[:, 1] + F(0.5) * h
    return x, y, w, h


def decode_coef(anchors, deltas, exp=None):
    """tube_utils.py:165-189.  `exp` lets the caller supply the exponential used by the device
    (fp32 expf is not bit-identical between libm, torch-CPU and CUDA); default np.exp in fp32."""
    anchors = np.asarray(anchors, dtype=F)
    deltas = np.asarray(deltas, dtype=F)
    exp = exp or (lambda v: np.exp(v.astype(F)).astype(F))
    x, y, w, h = get_center_size(anchors)
    px = w * deltas[:, 0] + x
    py = h * deltas[:, 1] + y
    pw = w * exp(deltas[:, 2])
    ph = h * exp(deltas[:, 3])
    out = np.empty_like(deltas)
    out[:, 0] = px - F(0.5) * pw
    out[:, 1] = py - F(0.5) * ph
    out[:, 2] = px + F(0.5) * pw - F(1.0)
    out[:, 3] = py + F(0.5) * ph - F(1.0)
    return out


def encode_coef(gt, tubes):
    """tube_utils.py:143-163."""
    gx, gy, gw, gh = get_center_size(gt)
    x, y, w, h = get_center_size(tubes)
    return np.stack(((gx - x) / w, (gy - y) / h, np.log(gw / w).astype(F), np.log(gh / h).astype(F)),
                    axis=1).astype(F)


def extrapolate_tubes(tubes, T, height=400, width=400):
    """tube_utils.py:10-27: linear recurrence outward from both ends, then clamp.
    NB the reference evaluates T/(T-1) and 1/(T-1) in Python double and multiplies fp32 arrays by
    them (numpy keeps fp32 for python-float scalars)."""
    tubes = np.asarray(tubes, dtype=F)
    n, L, d = tubes.shape
    out = np.zeros((n, L + 2 * T, d), dtype=F)
    out[:, T:-T] = tubes
    a = T / (T - 1)
    b = 1 / (T - 1)
    for i in range(T):
        out[:, -T + i] = a * out[:, -T + i - 1] - b * out[:, -T + i - T]
        out[:, T - i - 1] = a * out[:, T - i] - b * out[:, T - i + T - 1]
    out[:, :, 0] = np.maximum(0, out[:, :, 0])
    out[:, :, 1] = np.maximum(0, out[:, :, 1])
    out[:, :, 2] = np.minimum(width - 1, out[:, :, 2])
    out[:, :, 3] = np.minimum(height - 1, out[:, :, 3])
    return out


def valid_tubes(tubes, width=400, height=400):
    """tube_utils.py:59-92: clamp to [0,width]x[0,height]; a box failing x1<x2-2 and y1<y2-2
    becomes the whole image.  Works on a copy (the reference mutates in place)."""
    tubes = np.array(tubes, dtype=F, copy=True)
    n, T, _ = tubes.shape
    b = tubes.reshape(-1, 4)
    b[:, 0] = np.maximum(0, b[:, 0])
    b[:, 1] = np.maximum(0, b[:, 1])
    b[:, 2] = np.minimum(width, b[:, 2])
    b[:, 3] = np.minimum(height, b[:, 3])
    ok = (b[:, 0] < b[:, 2] - F(2)) & (b[:, 1] < b[:, 3] - F(2))
    b[~ok] = np.array([0, 0, width, height], dtype=F)
    return b.reshape(n, T, 4)


def flatten_tubes(tubes, batch_idx=False):
    """tube_utils.py:214-246: concat per-clip lists; column 0 = frame index arange(T)+i*T."""
    _, T, dim = tubes[0].shape
    flat, nums = [], []
    for i, t in enumerate(tubes):
        nums.append(t.shape[0])
        if t.shape[0] == 0:
            continue
        t = np.asarray(t, dtype=F)
        if batch_idx:
            idx = np.tile((np.arange(T) + i * T).reshape(1, T, 1), (t.shape[0], 1, 1)).astype(F)
            flat.append(np.concatenate((idx, t), axis=2))
        else:
            flat.append(t.copy())
    return np.concatenate(flat, axis=0), nums


def extend_tubes(tubes, ratio=1.2, width=400, height=400):
    """tube_utils.py:248-266 (dead code in the reference, named by north_star). tubes [-1,T,5]."""
    tubes = np.asarray(tubes, dtype=F)
    flat = tubes.reshape(-1, 5).copy()
    x, y, w, h = get_center_size(flat[:, 1:])
    w = w * F(ratio)
    h = h * F(ratio)
    flat[:, 1] = np.maximum(x - F(0.5) * w, 0)
    flat[:, 2] = np.maximum(y - F(0.5) * h, 0)
    flat[:, 3] = np.minimum(x + F(0.5) * w - F(1), F(width - 1))
    flat[:, 4] = np.minimum(y + F(0.5) * h - F(1), F(height - 1))
    return flat.reshape(tubes.shape)
