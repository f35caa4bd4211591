"""Shadow of `lib.model.DSTformer` (reference: lib/model/DSTformer.py).

The reference's `lib/` has no __init__.py (PEP 420 namespace package), so putting this directory
BEFORE the MotionBERT checkout on sys.path replaces exactly this one module while `lib.model.drop`,
`lib.utils.*`, `lib.data.*` still resolve from the reference:

    PYTHONPATH=/root/repo/shim:/root/repo:/path/to/MotionBERT python -P /path/to/MotionBERT/train.py ...

`lib/utils/learning.py::load_backbone` then builds the B200-native encoder with no source change.
"""
from motionbert_b200.dstformer import DSTformer  # noqa: F401

__all__ = ["DSTformer"]
