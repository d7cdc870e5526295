"""step_b200 -- the STEP (NVlabs/STEP) inference hot path on B200 (sm_100a).

Public surface mirrors the reference:
    from step_b200 import BaseNet, ROINet, TwoBranchNet, ContextNet      # models/__init__.py:6-7
    from step_b200 import inference                                       # utils/utils.py:15
    from step_b200.roi_layers import nms, roi_align, ROIAlign, roi_pool, ROIPool
    from step_b200 import tube_utils                                      # utils/tube_utils.py
All compute goes through libstep_b200.so (include/step_b200.h); there is no CPU/PyTorch fallback.
"""
from .networks import BaseNet, ROINet  # noqa: F401
from .two_branch import ContextNet, TwoBranchNet  # noqa: F401
from .inference import inference  # noqa: F401
from .runner import StepRunner  # noqa: F401
from . import postprocess, roi_layers, tube_utils  # noqa: F401

__all__ = ["BaseNet", "ROINet", "TwoBranchNet", "ContextNet", "inference", "StepRunner", "roi_layers", "tube_utils", "postprocess"]
