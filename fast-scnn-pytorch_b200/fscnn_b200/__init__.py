"""Host-side plumbing of the B200 Fast-SCNN forward path: the ctypes binding of
``csrc/libfscnn_b200.so`` (include/fscnn_b200.h) and the ``Engine`` that owns one native
context plus its PyTorch-allocated weight and workspace buffers."""
from .native import NativeError, lib, lib_path  # noqa: F401
from .engine import Engine, PRECISIONS  # noqa: F401
from .pipeline import StreamingEvaluator  # noqa: F401
from .trainer import Trainer, poly_lr  # noqa: F401
