"""Engine: one native context + its device buffers, all allocated by PyTorch.

The engine is what ``models.fast_scnn.FastSCNN`` (the drop-in nn.Module) delegates to in eval mode
on a CUDA tensor.  Ownership follows include/fscnn_b200.h: torch owns every buffer (packed
weights, workspace, outputs); the native library only enqueues kernels on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import native

PRECISIONS = {'fp32': native.PREC_FP32, 'bf16': native.PREC_BF16}
_DTYPE_CODE = {torch.uint8: native.U8, torch.int32: native.I32, torch.int64: native.I64}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    def __init__(self, num_classes: int, aux: bool, precision: str = 'fp32'):
        if precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {sorted(PRECISIONS)}, got {precision!r}')
        self.lib = native.lib()
        self.num_classes, self.aux, self.precision = int(num_classes), bool(aux), precision
        handle = C.c_void_p()
        native.check(self.lib.fscnn_create(C.byref(handle), self.num_classes, int(self.aux), PRECISIONS[precision]),
                     'fscnn_create')
        self._ctx = handle
        self._packed: Optional[torch.Tensor] = None
        self._ws: Dict[torch.device, torch.Tensor] = {}
        self.device: Optional[torch.device] = None
        self._input_state = (native.IN_F32_NCHW, None, None)

    def __del__(self):
        ctx, self._ctx = getattr(self, '_ctx', None), None
        if ctx:
            self.lib.fscnn_destroy(ctx)

    # ---- weights ---------------------------------------------------------------------------
    def param_manifest(self):
        n = self.lib.fscnn_param_count(self._ctx)
        return [(self.lib.fscnn_param_name(self._ctx, i).decode(), int(self.lib.fscnn_param_numel(self._ctx, i)))
                for i in range(n)]

    def load_state_dict(self, state_dict, device: torch.device) -> None:
        """Folds BN into the convolutions and repacks, on `device` (fscnn_load_weights)."""
        device = torch.device(device)
        manifest = self.param_manifest()
        keep, arr = [], (native.Tensor * len(manifest))()
        for i, (name, numel) in enumerate(manifest):
            if name not in state_dict:
                raise KeyError(f'state_dict is missing {name!r}')
            t = state_dict[name].detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            arr[i] = native.Tensor(name.encode(), t.data_ptr(), t.numel())
        nbytes = C.c_size_t()
        native.check(self.lib.fscnn_packed_weight_bytes(self._ctx, C.byref(nbytes)))
        with torch.cuda.device(device):
            packed = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
            native.check(self.lib.fscnn_load_weights(self._ctx, arr, len(manifest), packed.data_ptr(), nbytes.value,
                                                     _stream_ptr()), 'fscnn_load_weights')
            # the fold kernels read `keep` asynchronously; hold the staging tensors until they are done
            torch.cuda.current_stream().synchronize()
        self._packed, self.device = packed, device
        self._ws.clear()

    # ---- buffers ---------------------------------------------------------------------------
    def _workspace(self, n: int, h: int, w: int) -> torch.Tensor:
        need = C.c_size_t()
        native.check(self.lib.fscnn_workspace_bytes(self._ctx, n, h, w, C.byref(need)), 'fscnn_workspace_bytes')
        ws = self._ws.get(self.device)
        if ws is None or ws.numel() < need.value:
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            self._ws[self.device] = ws
        return ws

    def _set_input_format(self, fmt: int, mean, std) -> None:
        state = (fmt, tuple(mean) if mean is not None else None, tuple(std) if std is not None else None)
        if state == self._input_state:
            return
        f3 = C.c_float * 3
        native.check(self.lib.fscnn_set_input_format(self._ctx, fmt, f3(*state[1]) if state[1] else None,
                                                     f3(*state[2]) if state[2] else None), 'fscnn_set_input_format')
        self._input_state = state

    def _check_input(self, x: torch.Tensor, norm=None) -> Tuple[int, int, int]:
        """Accepts float32 [N,3,H,W] (the reference's normalised tensor) or raw uint8 [N,H,W,3] images; for
        the latter `norm` = (mean, std) is applied inside the stem kernel (None: /255 only)."""
        if self._packed is None:
            raise RuntimeError('Engine.load_state_dict has not been called')
        if not x.is_cuda or x.device != self.device:
            raise RuntimeError(f'input is on {x.device}, weights are on {self.device}; there is no CPU path')
        if not x.is_contiguous() or x.dim() != 4:
            raise ValueError('input must be a contiguous 4-d tensor')
        if x.dtype == torch.uint8:
            if x.shape[3] != 3:
                raise ValueError(f'uint8 input must be [N,H,W,3], got {tuple(x.shape)}')
            mean, std = norm if norm is not None else (None, None)
            self._set_input_format(native.IN_U8_NHWC, mean, std)
            return x.shape[0], x.shape[1], x.shape[2]
        if x.dtype != torch.float32 or x.shape[1] != 3:
            raise ValueError(f'expected float32 [N,3,H,W] or uint8 [N,H,W,3], got {x.dtype} {tuple(x.shape)}')
        self._set_input_format(native.IN_F32_NCHW, None, None)
        return x.shape[0], x.shape[2], x.shape[3]

    # ---- the forward path --------------------------------------------------------------------
    def forward_logits(self, x: torch.Tensor, want_aux: bool, norm=None):
        n, h, w = self._check_input(x, norm)
        with torch.cuda.device(self.device):
            ws = self._workspace(n, h, w)
            logits = torch.empty((n, self.num_classes, h, w), dtype=torch.float32, device=self.device)
            aux = torch.empty_like(logits) if (want_aux and self.aux) else None
            native.check(self.lib.fscnn_forward_logits(self._ctx, x.data_ptr(), n, h, w, logits.data_ptr(),
                                                       aux.data_ptr() if aux is not None else None, ws.data_ptr(),
                                                       ws.numel(), _stream_ptr()), 'fscnn_forward_logits')
        return logits, aux

    def _check_class_map(self, t: torch.Tensor, n: int, h: int, w: int, what: str) -> None:
        """A caller-provided class map goes to the kernels as a raw pointer: shape, layout, device, dtype and the
        4-element alignment the vectorised accesses need (include/fscnn_b200.h) are checked here."""
        if t.dtype not in _DTYPE_CODE:
            raise ValueError(f'{what} must be torch.uint8, torch.int32 or torch.int64, got {t.dtype}')
        if tuple(t.shape) != (n, h, w) or not t.is_contiguous():
            raise ValueError(f'{what} must be a contiguous [{n},{h},{w}] tensor, got {tuple(t.shape)}')
        if t.device != self.device:
            raise RuntimeError(f'{what} is on {t.device}, the model is on {self.device}')
        align = 4 if t.dtype == torch.uint8 else 16           # uchar4 / int4 / longlong2 accesses
        if t.data_ptr() % align:
            raise ValueError(f'{what} must be {align}-byte aligned; pass a tensor that starts at such an offset of its storage')

    def forward_mask(self, x: torch.Tensor, out_dtype=torch.uint8, out: Optional[torch.Tensor] = None, norm=None) -> torch.Tensor:
        n, h, w = self._check_input(x, norm)
        if out_dtype not in _DTYPE_CODE:
            raise ValueError('mask dtype must be torch.uint8, torch.int32 or torch.int64')
        if out is not None:
            self._check_class_map(out, n, h, w, 'out')
        with torch.cuda.device(self.device):
            ws = self._workspace(n, h, w)
            mask = out if out is not None else torch.empty((n, h, w), dtype=out_dtype, device=self.device)
            native.check(self.lib.fscnn_forward_mask(self._ctx, x.data_ptr(), n, h, w, mask.data_ptr(),
                                                     _DTYPE_CODE[mask.dtype], ws.data_ptr(), ws.numel(), _stream_ptr()),
                         'fscnn_forward_mask')
        return mask

    def forward_confusion(self, x: torch.Tensor, labels: torch.Tensor, conf: torch.Tensor,
                          mask: Optional[torch.Tensor] = None, norm=None) -> torch.Tensor:
        n, h, w = self._check_input(x, norm)
        self._check_class_map(labels, n, h, w, 'labels')
        if mask is not None:
            self._check_class_map(mask, n, h, w, 'mask')
        if conf.device != self.device:
            raise RuntimeError('conf must be on the model device')
        if conf.dtype != torch.int64 or conf.numel() != self.conf_len() or not conf.is_contiguous():
            raise ValueError(f'conf must be a contiguous int64[{self.conf_len()}] tensor')
        with torch.cuda.device(self.device):
            ws = self._workspace(n, h, w)
            native.check(self.lib.fscnn_forward_confusion(
                self._ctx, x.data_ptr(), labels.data_ptr(), _DTYPE_CODE[labels.dtype], n, h, w, conf.data_ptr(),
                mask.data_ptr() if mask is not None else None, _DTYPE_CODE[mask.dtype] if mask is not None else 0,
                ws.data_ptr(), ws.numel(), _stream_ptr()), 'fscnn_forward_confusion')
        return conf

    def upsample_argmax(self, low_logits: torch.Tensor, h: int, w: int, out_dtype=torch.uint8, labels: Optional[torch.Tensor] = None,
                        conf: Optional[torch.Tensor] = None, want_mask: bool = True, exhaustive: bool = False) -> Optional[torch.Tensor]:
        """The tail stage alone (fscnn_upsample_argmax): low_logits float32 [N,hl,wl,padded_classes] (NHWC, the layout of the
        'cls.logits_lowres' tap) -> class map [N,h,w]; with labels, also adds the confusion counts into conf.  exhaustive=True
        switches the exact class pruning off (same result, worst-case time)."""
        if low_logits.dtype != torch.float32 or low_logits.dim() != 4 or not low_logits.is_contiguous() or not low_logits.is_cuda:
            raise ValueError('low_logits must be a contiguous CUDA float32 [N,hl,wl,padded_classes] tensor')
        n, hl, wl, ncp = low_logits.shape
        dev = low_logits.device
        mask = torch.empty((n, h, w), dtype=out_dtype, device=dev) if want_mask else None
        if labels is not None:
            if conf is None or conf.dtype != torch.int64 or conf.numel() != self.conf_len() or conf.device != dev:
                raise ValueError(f'labels need conf: an int64[{self.conf_len()}] tensor on {dev}')
            if labels.dtype not in _DTYPE_CODE or tuple(labels.shape) != (n, h, w) or not labels.is_contiguous() or labels.device != dev:
                raise ValueError('labels must be a contiguous [N,h,w] uint8/int32/int64 tensor on the logits\' device')
        with torch.cuda.device(dev):
            native.check(self.lib.fscnn_upsample_argmax(
                low_logits.data_ptr(), self.num_classes, ncp, n, hl, wl, h, w, mask.data_ptr() if mask is not None else None,
                _DTYPE_CODE[out_dtype], labels.data_ptr() if labels is not None else None,
                _DTYPE_CODE[labels.dtype] if labels is not None else 0, conf.data_ptr() if conf is not None else None,
                native.TAIL_EXHAUSTIVE if exhaustive else 0, _stream_ptr()), 'fscnn_upsample_argmax')
        return mask

    def conf_len(self) -> int:
        return int(self.lib.fscnn_conf_len(self.num_classes))

    # ---- test / profiling hooks ----------------------------------------------------------------
    def stage_names(self):
        return [self.lib.fscnn_stage_name(self._ctx, i).decode() for i in range(self.lib.fscnn_stage_count(self._ctx))]

    def forward_range(self, x: torch.Tensor, first: int, last: int, norm=None) -> None:
        n, h, w = self._check_input(x, norm)
        with torch.cuda.device(self.device):
            ws = self._workspace(n, h, w)
            native.check(self.lib.fscnn_forward_range(self._ctx, x.data_ptr(), n, h, w, first, last, ws.data_ptr(),
                                                      ws.numel(), _stream_ptr()), 'fscnn_forward_range')

    def tap_view(self, name: str, n: int, h: int, w: int) -> torch.Tensor:
        """A [n,h',w',c] view (NHWC) of a stage tensor inside the workspace."""
        tap = native.Tap()
        native.check(self.lib.fscnn_tap_info(self._ctx, n, h, w, name.encode(), C.byref(tap)), 'fscnn_tap_info')
        ws = self._workspace(n, h, w)
        dtype = torch.float32 if tap.elem_bytes == 4 else torch.bfloat16
        count = tap.n * tap.h * tap.w * tap.c_stride
        flat = ws[tap.offset_bytes: tap.offset_bytes + count * tap.elem_bytes].view(dtype)
        return flat.view(tap.n, tap.h, tap.w, tap.c_stride)[..., :tap.c]

    # ---- camera-frame wrapper (reference export_onnx_fixed.py:34-98) ---------------------
    def e2e_preprocess(self, frames: torch.Tensor, base_size: int, mean=None, std=None) -> torch.Tensor:
        """frames [N,3,h,w] uint8 / float32 (0..255) -> float32 [N,3,base,base]: resize (align_corners=False), /255, normalise."""
        if frames.dim() != 4 or frames.shape[1] != 3 or frames.dtype not in (torch.uint8, torch.float32):
            raise ValueError(f'expected uint8 or float32 frames [N,3,H,W], got {frames.dtype} {tuple(frames.shape)}')
        frames = frames.contiguous()
        n, _, h, w = frames.shape
        out = torch.empty((n, 3, base_size, base_size), dtype=torch.float32, device=frames.device)
        m = (C.c_float * 3)(*[float(v) for v in mean]) if mean is not None else None
        sd = (C.c_float * 3)(*[float(v) for v in std]) if std is not None else None
        with torch.cuda.device(frames.device):
            native.check(self.lib.fscnn_e2e_preprocess(frames.data_ptr(), native.U8 if frames.dtype == torch.uint8 else native.F32, n, h, w,
                                                       base_size, m, sd, out.data_ptr(), _stream_ptr()), 'fscnn_e2e_preprocess')
        return out

    def e2e_forward(self, x: torch.Tensor, out_h: int, out_w: int, apply_softmax: bool, chunk: int = 32) -> torch.Tensor:
        """network on x [N,3,H,W] float32 up to its low-resolution logits, then the fused x8 upsample + resize to (out_h, out_w)
        (+ softmax): float32 [N,nc,out_h,out_w].  The full-resolution logits are never materialised."""
        n, h, w = self._check_input(x, None)
        names = self.stage_names()
        last = names.index('cls.dsconv2+head')
        out = torch.empty((n, self.num_classes, out_h, out_w), dtype=torch.float32, device=self.device)
        probe = native.Tap()      # fscnn_forward_range takes at most one micro-batch: chunk by what the engine plans for this shape
        native.check(self.lib.fscnn_tap_info(self._ctx, n, h, w, b'cls.logits_lowres', C.byref(probe)), 'fscnn_tap_info')
        chunk = max(1, min(chunk, probe.n))
        with torch.cuda.device(self.device):
            for i0 in range(0, n, chunk):
                xs = x[i0:i0 + chunk]
                m = xs.shape[0]
                self.forward_range(xs, 0, last)
                tap = native.Tap()
                native.check(self.lib.fscnn_tap_info(self._ctx, m, h, w, b'cls.logits_lowres', C.byref(tap)), 'fscnn_tap_info')
                ws = self._workspace(m, h, w)
                native.check(self.lib.fscnn_e2e_postprocess(ws.data_ptr() + tap.offset_bytes, self.num_classes, tap.c_stride, m, tap.h, tap.w,
                                                            h, w, out_h, out_w, int(bool(apply_softmax)), out[i0:i0 + m].data_ptr(),
                                                            _stream_ptr()), 'fscnn_e2e_postprocess')
        return out

    def launch_count(self) -> int:
        return int(self.lib.fscnn_launch_count(self._ctx))

    def set_option(self, key: str, value: int) -> None:
        native.check(self.lib.fscnn_set_option(self._ctx, key.encode(), int(value)), 'fscnn_set_option')
        self._ws.clear()

    def set_micro_batch(self, images: int) -> None:
        native.check(self.lib.fscnn_set_micro_batch(self._ctx, int(images)))
        self._ws.clear()
