"""motionbert_b200 -- B200-native (sm_100a) DSTformer encoder behind MotionBERT's own class boundary.

    from motionbert_b200 import DSTformer            # same constructor / state_dict as lib.model.DSTformer
    PYTHONPATH=/root/repo/shim:/path/to/MotionBERT python -P train.py ...   # reference scripts unchanged

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of
include/motionbert_b200.h), the ctypes binding (`_lib`), the host-side mirror of the reference
class (`dstformer`, with `_autograd` wiring `loss.backward()` to the native backward), and the two
steps either side of the path: `loss` (pretrain losses fused with their gradient), `tta` (flip-TTA as
one call), `dist` (sharding + gradient exchange).  Importing it never builds anything;
`python -m motionbert_b200.build` does.
"""
from .dstformer import DSTformer  # noqa: F401

__all__ = ["DSTformer"]
