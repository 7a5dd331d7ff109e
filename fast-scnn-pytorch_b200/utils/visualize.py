"""Palette output for class maps, with the reference's interface (utils/visualize.py:7-36) plus a device-side
renderer for the masks the fused kernels produce.

``get_color_pallete(npimg, dataset='citys')`` keeps the reference's contract: a numpy class map in, a PIL 'P'-mode
image carrying the dataset palette out (what eval.py:53-55 / demo.py:49-51 save as PNG).  ``colorize(mask, dataset)``
is the addition: it maps a CUDA class-map tensor [N,H,W] (uint8 / int32 / int64) to an RGB uint8 tensor [N,H,W,3] with one
kernel (``fscnn_colorize``), so only the final picture crosses PCIe.

Palettes: 'citys' = the 19 Cityscapes train-id colours; anything else = the Pascal-VOC bit-interleaved colour map
(the reference's default branch).  The reference's 150-colour ADE20K table belongs to a dataset this repository's
models are never built for (data_loader/__init__.py:6-11) and is not reproduced.
"""
from __future__ import annotations

import numpy as np

__all__ = ['get_color_pallete', 'colorize', 'overlay', 'palette_for']

_CITYS = (128, 64, 128, 244, 35, 232, 70, 70, 70, 102, 102, 156, 190, 153, 153, 153, 153, 153, 250, 170, 30, 220, 220, 0,
          107, 142, 35, 152, 251, 152, 0, 130, 180, 220, 20, 60, 255, 0, 0, 0, 0, 142, 0, 0, 70, 0, 60, 100, 0, 80, 100,
          0, 0, 230, 119, 11, 32)


def _voc_palette(n=256):
    """Colour j takes bit b of j into bit (7 - b//3) of channel b%3 -- the standard VOC colour map."""
    pal = np.zeros((n, 3), dtype=np.uint8)
    for j in range(n):
        for b in range(24):
            if (j >> b) & 1:
                pal[j, b % 3] |= 1 << (7 - b // 3)
    return pal


_VOC = _voc_palette()


def palette_for(dataset='citys') -> np.ndarray:
    """uint8 [256, 3] palette (unused entries are black for 'citys')."""
    if dataset == 'ade20k':
        raise ValueError("the ADE20K palette is not part of this build (no model of this repository uses that dataset)")
    if dataset == 'citys':
        pal = np.zeros((256, 3), dtype=np.uint8)
        pal[:19] = np.asarray(_CITYS, dtype=np.uint8).reshape(19, 3)
        return pal
    return _VOC.copy()


def get_color_pallete(npimg, dataset='citys'):
    """Class map (numpy, H x W) -> PIL 'P' image with the dataset palette."""
    from PIL import Image
    npimg = np.asarray(npimg)
    if dataset in ('pascal_voc', 'pascal_aug'):
        npimg = np.where(npimg == -1, 255, npimg)       # recover the boundary label
    out_img = Image.fromarray(npimg.astype('uint8'))
    pal = palette_for(dataset)
    out_img.putpalette(pal[:19].reshape(-1).tolist() if dataset == 'citys' else pal.reshape(-1).tolist())
    return out_img


def colorize(mask, dataset='citys', out=None):
    """CUDA class-map tensor [...,H,W] -> uint8 RGB tensor [...,H,W,3] (rgb = palette[mask & 255])."""
    import torch
    from fscnn_b200 import native
    codes = {torch.uint8: native.U8, torch.int32: native.I32, torch.int64: native.I64}
    if not mask.is_cuda or mask.dtype not in codes:
        raise ValueError('colorize expects a CUDA uint8 / int32 / int64 class map (use get_color_pallete for numpy arrays)')
    mask = mask.contiguous()
    rgb = out if out is not None else torch.empty(tuple(mask.shape) + (3,), dtype=torch.uint8, device=mask.device)
    pal = palette_for(dataset).tobytes()
    with torch.cuda.device(mask.device):
        native.check(native.lib().fscnn_colorize(mask.data_ptr(), codes[mask.dtype], mask.numel(), pal, rgb.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), 'fscnn_colorize')
    return rgb


def overlay(image, mask, classes=(1,), colors=None, dataset='tusimple', alpha=0.5, out=None):
    """Blend the pixels of ``classes`` towards their colour on the frame (reference demo_tusimple.py:87-104 ``create_overlay``:
    ``overlay[m] = (1 - alpha) * image[m] + alpha * colour``, truncated to uint8), on the device.

    image: CUDA uint8 [...,H,W,3]; mask: CUDA class map [...,H,W] (uint8 / int32 / int64); ``colors``: {class: (r, g, b)} overriding
    the dataset palette (the reference draws lane class 1 in green: ``overlay(img, mask, classes=(1,), colors={1: (0, 255, 0)})``)."""
    import ctypes as C
    import torch
    from fscnn_b200 import native
    codes = {torch.uint8: native.U8, torch.int32: native.I32, torch.int64: native.I64}
    if not (image.is_cuda and mask.is_cuda) or image.dtype != torch.uint8 or mask.dtype not in codes:
        raise ValueError('overlay expects a CUDA uint8 [...,H,W,3] frame and a CUDA uint8 / int32 / int64 class map')
    if tuple(image.shape) != tuple(mask.shape) + (3,) or image.device != mask.device:
        raise ValueError(f'frame {tuple(image.shape)} and class map {tuple(mask.shape)} do not match')
    image, mask = image.contiguous(), mask.contiguous()
    pal = palette_for(dataset)
    for c, rgb in (colors or {}).items():
        pal[int(c)] = rgb
    draw = (C.c_uint * 8)()
    for c in classes:
        if not 0 <= int(c) < 256:
            raise ValueError(f'class {c} outside [0, 256)')
        draw[int(c) >> 5] |= 1 << (int(c) & 31)
    res = out if out is not None else torch.empty_like(image)
    with torch.cuda.device(image.device):
        native.check(native.lib().fscnn_overlay(image.data_ptr(), mask.data_ptr(), codes[mask.dtype], mask.numel(), pal.tobytes(), draw,
                                                float(alpha), res.data_ptr(), torch.cuda.current_stream().cuda_stream), 'fscnn_overlay')
    return res
