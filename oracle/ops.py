"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/step_oracle.c and loader of the
reference's own compiled CPU ops (oracle/_ref/_C.so, see oracle/build_ref.py).

All functions take/return numpy arrays (fp32 / int64 / int32).
"""
import ctypes
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "libstep_oracle.so")
        if not os.path.exists(path):
            from oracle import build_ref
            build_ref.build_c_oracle()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_nms_f32.restype = ctypes.c_int64
    return _LIB


def ref_C():
    """The reference's own `_C` module (nms, roi_align_forward) or None if it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "_C.so")
        if not os.path.exists(path):
            return None
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location("_C", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF = mod
    return _REF


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nms(boxes, scores, thr, ge=True):
    """cpu/nms_cpu.cpp:29-89 (ge=True) / cuda/nms.cu (ge=False). Returns int64 kept indices ascending."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().orc_nms_f32(_p(boxes), _p(scores), ctypes.c_int64(n), ctypes.c_float(thr),
                          ctypes.c_int(1 if ge else 0), _p(keep))
    return keep[:k].copy()


def roi_align_fwd(feat, rois, scale, ph, pw, sampling_ratio):
    """cpu/ROIAlign_cpu.cpp:137-243. feat [K,C,H,W] fp32, rois [R,5] -> [R,C,ph,pw]."""
    feat = _f32(feat)
    rois = _f32(rois).reshape(-1, 5)
    K, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), dtype=np.float32)
    lib().orc_roi_align_fwd_f32(_p(feat), K, C, H, W, _p(rois), R, ctypes.c_float(scale), ph, pw,
                                sampling_ratio, _p(out))
    return out


def roi_align_bwd(grad_out, rois, scale, ph, pw, K, C, H, W, sampling_ratio):
    """cuda/ROIAlign_cuda.cu:201-278 restated sequentially."""
    grad_out = _f32(grad_out)
    rois = _f32(rois).reshape(-1, 5)
    R = rois.shape[0]
    gin = np.empty((K, C, H, W), dtype=np.float32)
    lib().orc_roi_align_bwd_f32(_p(grad_out), _p(rois), R, ctypes.c_float(scale), ph, pw, K, C, H, W,
                                sampling_ratio, _p(gin))
    return gin


def roi_pool_fwd(feat, rois, scale, ph, pw):
    """cuda/ROIPool_cuda.cu:40-101. Returns (out fp32, argmax int32)."""
    feat = _f32(feat)
    rois = _f32(rois).reshape(-1, 5)
    K, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), dtype=np.float32)
    arg = np.empty((R, C, ph, pw), dtype=np.int32)
    lib().orc_roi_pool_fwd_f32(_p(feat), K, C, H, W, _p(rois), R, ctypes.c_float(scale), ph, pw,
                               _p(out), _p(arg))
    return out, arg


def roi_pool_bwd(grad_out, argmax, rois, ph, pw, K, C, H, W):
    """cuda/ROIPool_cuda.cu:103-132."""
    grad_out = _f32(grad_out)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    rois = _f32(rois).reshape(-1, 5)
    R = rois.shape[0]
    gin = np.empty((K, C, H, W), dtype=np.float32)
    lib().orc_roi_pool_bwd_f32(_p(grad_out), _p(argmax), _p(rois), R, ph, pw, K, C, H, W, _p(gin))
    return gin
