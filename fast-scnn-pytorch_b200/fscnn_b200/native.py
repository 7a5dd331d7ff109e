"""ctypes binding of libfscnn_b200.so -- one Python declaration per symbol of
include/fscnn_b200.h.  There is no fallback: if the shared library is missing the import of the
forward path fails loudly (build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C fast-scnn-pytorch_b200/csrc``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = 'libfscnn_b200.so'

PREC_FP32, PREC_BF16 = 0, 1
U8, I32, I64, F32 = 0, 1, 2, 3
IN_F32_NCHW, IN_U8_NHWC = 0, 1
TAIL_EXHAUSTIVE = 1


class NativeError(RuntimeError):
    """A call into libfscnn_b200.so returned a negative status."""


class Tensor(C.Structure):
    _fields_ = [('name', C.c_char_p), ('d_data', C.c_void_p), ('numel', C.c_int64)]


class Tap(C.Structure):
    _fields_ = [('offset_bytes', C.c_size_t), ('n', C.c_int), ('h', C.c_int), ('w', C.c_int), ('c', C.c_int),
                ('c_stride', C.c_int), ('elem_bytes', C.c_int)]


def lib_path() -> str:
    return os.environ.get('FSCNN_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'csrc', _LIB_NAME)


_SIGNATURES = {
    # name: (restype, argtypes)
    'fscnn_abi_version': (C.c_int, []),
    'fscnn_last_error': (C.c_char_p, []),
    'fscnn_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    'fscnn_destroy': (None, [C.c_void_p]),
    'fscnn_param_count': (C.c_int, [C.c_void_p]),
    'fscnn_param_name': (C.c_char_p, [C.c_void_p, C.c_int]),
    'fscnn_param_numel': (C.c_int64, [C.c_void_p, C.c_int]),
    'fscnn_packed_weight_bytes': (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    'fscnn_load_weights': (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'fscnn_set_input_format': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    'fscnn_workspace_bytes': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'fscnn_forward_logits': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    'fscnn_forward_mask': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    'fscnn_forward_confusion': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'fscnn_conf_len': (C.c_int64, [C.c_int]),
    'fscnn_confusion_from_mask': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                            C.c_void_p]),
    'fscnn_colorize': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_char_p, C.c_void_p, C.c_void_p]),
    'fscnn_overlay': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_char_p, C.POINTER(C.c_uint), C.c_double, C.c_void_p, C.c_void_p]),
    'fscnn_e2e_preprocess': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.c_void_p, C.c_void_p]),
    'fscnn_e2e_postprocess': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    'fscnn_upsample_argmax': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'fscnn_conf_to_totals': (C.c_int, [C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_longlong),
                                       C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    'fscnn_train_workspace_bytes': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'fscnn_train_dwconv3x3_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_dwconv3x3_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_pwconv_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_pwconv_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_batchnorm_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                C.c_int, C.c_void_p]),
    'fscnn_train_batchnorm_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p]),
    'fscnn_train_stem_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_stem_weight_grad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_im2col3x3': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_col2im3x3': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_bias_add': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_bias_grad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_bilinear': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_adaptive_avg_pool': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'fscnn_train_dropout': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_ulonglong, C.c_void_p, C.c_int64, C.c_void_p]),
    'fscnn_train_add_relu': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    'fscnn_train_relu_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'fscnn_train_set_math': (C.c_int, [C.c_int]),
    'fscnn_train_get_math': (C.c_int, []),
    'fscnn_train_sgd_step': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int64,
                                       C.c_void_p]),
    'fscnn_train_adamw_step': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_double, C.c_double, C.c_float, C.c_float,
                                         C.c_float, C.c_int64, C.c_int64, C.c_void_p]),
    'fscnn_train_ohem_workspace_bytes': (C.c_int, [C.POINTER(C.c_size_t)]),
    'fscnn_train_ohem_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                           C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_int, C.c_void_p]),
    'fscnn_train_ohem_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]),
    'fscnn_train_ohem_upsampled_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_int,
                                                     C.c_void_p]),
    'fscnn_train_ohem_upsampled_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                                      C.c_void_p]),
    'fscnn_train_criterion_workspace_bytes': (C.c_int, [C.POINTER(C.c_size_t)]),
    'fscnn_train_criterion_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    'fscnn_train_criterion_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    'fscnn_stage_count': (C.c_int, [C.c_void_p]),
    'fscnn_stage_name': (C.c_char_p, [C.c_void_p, C.c_int]),
    'fscnn_forward_range': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    'fscnn_tap_info': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(Tap)]),
    'fscnn_launch_count': (C.c_int64, [C.c_void_p]),
    'fscnn_set_micro_batch': (C.c_int, [C.c_void_p, C.c_int]),
    'fscnn_set_option': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def lib():
    """The loaded shared library with typed entry points (loaded once)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(f'{path} not found: the CUDA extension is not built (make -C '
                              f'{os.path.dirname(path)}); there is no CPU fallback for this path')
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.fscnn_abi_version() != 1:
            raise ImportError(f'{path}: ABI version {handle.fscnn_abi_version()} != 1')
        _lib = handle
    return _lib


def check(status: int, what: str = '') -> None:
    if status != 0:
        msg = lib().fscnn_last_error().decode('utf-8', 'replace')
        raise NativeError(f'{what or "libfscnn_b200"} failed ({status}): {msg}')
