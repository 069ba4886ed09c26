"""CPU oracle (test infrastructure only) for the frame pre-processing that feeds the hot path
(SURVEY.md §8f-3):  video `.pt` tensor / decoded frames -> SigLipImageProcessor.preprocess -> pixel_values.

Reference call sites (never imported by the product path):
  llava/model/multimodal_encoder/siglip_encoder.py:34-67   SigLipImageProcessor: convert_to_rgb -> to_numpy_array ->
      resize(size=(384, 384), resample=PILImageResampling.BICUBIC) -> rescale(1/255) -> normalize(mean 0.5, std 0.5)
      -> channels-first
  llava/train/train.py:1230-1241                            video tensor -> processor.preprocess(video)["pixel_values"]
  extract_video_frames/video_reader_tmp.py:82-88            the `.pt` file: decord frames, uint8 [F, H, W, 3]

The arithmetic lives in third-party code that is not under /root/reference:
  * Pillow (pinned by the reference: requirements.txt `pillow==10.2.0`; this container has 12.2.0; the resampling
    code is unchanged between them), src/libImaging/Resample.c: `precompute_coeffs`, `normalize_coeffs_8bpc`,
    `ImagingResampleHorizontal_8bpc` / `Vertical_8bpc` -- a separable two-pass filter in 22-bit fixed point with a
    uint8 intermediate image, horizontal pass first, a pass skipped when that dimension already has the target
    size.  Restated below from the published algorithm.
  * transformers.image_transforms.rescale / normalize: float64(u8) * scale -> float32, then (x - mean) / std in
    float32.
Pinned in tests/test_preprocess_oracle.py against PIL + transformers executed in this container.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BICUBIC_SUPPORT = 2.0


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter (a = -0.5)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).
    Returns (ksize, bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            v = w * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w < 0 else int(0.5 + v)      # C (int) cast truncates toward zero
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One 8bpc pass along `axis` of a uint8 [..., H, W, C] image."""
    out_size = bounds.shape[0]
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)      # clip8
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(frames: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BICUBIC) for uint8 frames [F, H, W, C] (RGB)."""
    assert frames.dtype == np.uint8 and frames.ndim == 4
    f, h, w, c = frames.shape
    img = frames
    if w != out_w:                                            # horizontal pass first (Resample.c ImagingResample)
        _, bh, kh = precompute_coeffs(w, out_w)
        img = _pass(img, bh, kh, axis=2)
    if h != out_h:
        _, bv, kv = precompute_coeffs(h, out_h)
        img = _pass(img, bv, kv, axis=1)
    return np.ascontiguousarray(img)


def rescale_normalize(u8: np.ndarray, rescale_factor=1 / 255, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """transformers rescale + normalize + channels-first: uint8 [F, H, W, 3] -> float32 [F, 3, H, W]."""
    x = (u8.astype(np.float64) * rescale_factor).astype(np.float32)
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def preprocess(frames: np.ndarray, size=(384, 384), rescale_factor=1 / 255, mean=(0.5, 0.5, 0.5),
               std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """SigLipImageProcessor.preprocess(...)["pixel_values"] for a uint8 video tensor [F, H, W, 3]."""
    return rescale_normalize(resize_bicubic_u8(frames, size[0], size[1]), rescale_factor, mean, std)
