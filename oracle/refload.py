"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference Python (models/, utils/,
external/maskrcnn_benchmark/roi_layers) from /root/reference, wiring in the reference's own CPU ops
compiled by oracle/build_ref.py.  Exists only in the build container (the GPU box has no
/root/reference): callers must check `available()` and skip otherwise.  Used to pin oracle/ and to
generate tests/golden/*.npz (tests/golden/make_golden.py)."""
import os
import sys

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def load():
    """Returns a namespace with the reference symbols on the hot path."""
    from types import SimpleNamespace
    from . import build_ref, ops
    assert available(), "reference tree not present"
    build_ref.build_reference_ops()
    C = ops.ref_C()
    sys.modules.setdefault("external.maskrcnn_benchmark.roi_layers._C", C)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # apex warning print at import
        import models as ref_models
        from utils import tube_utils as ref_tubes
        from utils import utils as ref_utils
        from external.maskrcnn_benchmark import roi_layers as ref_roi
    return SimpleNamespace(models=ref_models, tube_utils=ref_tubes, utils=ref_utils, roi_layers=ref_roi, _C=C)
