"""Frame pre-processing on the GPU: the producer side of the visual-memory path.

`SigLipImageProcessor` mirrors llava/model/multimodal_encoder/siglip_encoder.py:34-67 -- same constructor
arguments, same `preprocess(images, return_tensors)` call, same `{"pixel_values": ...}` result -- but runs the
resize (PIL bicubic, bit-exact), rescale, normalize and HWC->CHW on the device in two kernels of libmavlm.so
(csrc/preprocess.cu).  Input: a uint8 video tensor [F, H, W, 3] (what extract_video_frames/video_reader_tmp.py
stores and train.py:1234 loads), a numpy array of that shape, or a list of equally sized HxWx3 uint8 frames.
The reference's per-frame host loop (PIL + numpy, ~8 ms per 720p frame on one core) becomes ~1 us per frame
of HBM traffic; the frames cross PCIe as uint8 (3 B / pixel) instead of float32 pixel_values (12 B / pixel).
No CPU fallback: CPU tensors are moved to the current CUDA device (that is the H2D copy), never processed on the host.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._device import on_tensor_device
from .ops import _DTYPES, _ptr, _stream


class BatchFeature(dict):
    """Minimal stand-in for transformers.BatchFeature (dict with attribute access), siglip_encoder.py:67."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item) from None


_TABLES: Dict[Tuple[int, int, str], Tuple[int, torch.Tensor, torch.Tensor]] = {}


def resize_tables(in_size: int, out_size: int, device) -> Tuple[int, torch.Tensor, torch.Tensor]:
    """(ksize, bounds int32 [out, 2], kk int32 [out, ksize]) of Pillow's bicubic resampler for one axis, on `device`."""
    key = (in_size, out_size, str(device))
    if key not in _TABLES:
        lib = _lib.load()
        ksize = lib.mavlm_resize_coeffs(in_size, out_size, None, None, 0)
        if ksize < 0:
            _lib.check(ksize, "resize_coeffs")
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        kk = np.zeros((out_size, ksize), dtype=np.int32)
        st = lib.mavlm_resize_coeffs(in_size, out_size, bounds.ctypes.data, kk.ctypes.data, ksize)
        if st < 0:
            _lib.check(st, "resize_coeffs")
        _TABLES[key] = (ksize, torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device))
    return _TABLES[key]


@on_tensor_device
def frames_preprocess(frames: torch.Tensor, size: Tuple[int, int] = (384, 384), rescale_factor: float = 1 / 255,
                      image_mean: Sequence[float] = (0.5, 0.5, 0.5), image_std: Sequence[float] = (0.5, 0.5, 0.5),
                      dtype: torch.dtype = torch.float32, return_resized: bool = False):
    """uint8 CUDA frames [F, H, W, 3] -> pixel_values [F, 3, size[0], size[1]] (`dtype` float32 or bfloat16)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected uint8 frames [F, H, W, 3], got {frames.dtype} {tuple(frames.shape)}")
    if not frames.is_cuda:
        raise RuntimeError("mavlm: frames are not on a CUDA device; this path has no CPU fallback")
    if len(image_mean) != 3 or len(image_std) != 3:
        raise ValueError("mean / std must have 3 elements")                   # transformers.normalize's check
    frames = frames.contiguous()
    f, h, w, _ = frames.shape
    oh, ow = int(size[0]), int(size[1])
    dev = frames.device
    kh = kv = 0
    bh = kkh = bv = kkv = None
    tmp = None
    if w != ow:
        kh, bh, kkh = resize_tables(w, ow, dev)
        tmp = torch.empty((f, h, ow, 3), dtype=torch.uint8, device=dev)
    if h != oh:
        kv, bv, kkv = resize_tables(h, oh, dev)
    out = torch.empty((f, 3, oh, ow), dtype=dtype, device=dev)
    u8 = torch.empty((f, oh, ow, 3), dtype=torch.uint8, device=dev) if return_resized else None
    mean = (ctypes.c_float * 3)(*[float(m) for m in image_mean])
    std = (ctypes.c_float * 3)(*[float(s) for s in image_std])
    st = _lib.load().mavlm_frames_preprocess_fwd(_ptr(frames), f, h, w, _ptr(tmp), _ptr(u8), _ptr(out), oh, ow, _ptr(bh),
                                                 _ptr(kkh), kh, _ptr(bv), _ptr(kkv), kv, float(rescale_factor),
                                                 ctypes.addressof(mean), ctypes.addressof(std), _DTYPES[dtype], _stream())
    _lib.check(st, "frames_preprocess_fwd")
    return (out, u8) if return_resized else out


class SigLipImageProcessor:
    """Drop-in for siglip_encoder.py:34-67 (resample is fixed to BICUBIC = 3, data_format to channels-first: the only
    values the reference ever constructs it with, siglip_encoder.py:35, 548)."""

    def __init__(self, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), size=(384, 384),
                 crop_size: Optional[Dict[str, int]] = None, resample=3, rescale_factor=1 / 255, data_format="channels_first",
                 device: Union[str, torch.device, None] = None, dtype: torch.dtype = torch.float32):
        if int(resample) != 3:
            raise ValueError("only PILImageResampling.BICUBIC (3) is implemented on this path")
        if str(getattr(data_format, "value", data_format)) != "channels_first":
            raise ValueError("only channels-first output is implemented on this path")
        self.image_mean = image_mean
        self.image_std = image_std
        self.size = size
        self.resample = resample
        self.rescale_factor = rescale_factor
        self.data_format = data_format
        self.crop_size = crop_size if crop_size is not None else {"height": 384, "width": 384}
        self.device = device
        self.dtype = dtype

    def _to_device_u8(self, images) -> torch.Tensor:
        if isinstance(images, torch.Tensor):
            t = images
        elif isinstance(images, np.ndarray):
            t = torch.from_numpy(images)
        else:                                                       # list of frames (numpy / PIL / tensors)
            t = torch.from_numpy(np.stack([np.asarray(im.convert("RGB")) if hasattr(im, "convert") else np.asarray(im)
                                           for im in images]))
        if t.dim() == 3:
            t = t[None]
        dev = self.device or (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        return t.to(dev, non_blocking=True)

    def preprocess(self, images, return_tensors="pt"):
        if return_tensors not in ("pt", None):
            raise ValueError("this path returns torch tensors (return_tensors='pt')")
        x = frames_preprocess(self._to_device_u8(images), self.size, self.rescale_factor, self.image_mean,
                              self.image_std, self.dtype)
        return BatchFeature(pixel_values=x)
