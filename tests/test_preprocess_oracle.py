"""The pre-processing oracle against the real third-party code the reference calls (PIL + transformers), run here."""
import numpy as np
import pytest

from oracle import preprocess_oracle as po

PIL = pytest.importorskip("PIL.Image")


def _pil_resize(frames, h, w):
    from PIL import Image
    return np.stack([np.array(Image.fromarray(f).resize((w, h), resample=Image.BICUBIC)) for f in frames])


@pytest.mark.parametrize("shape", [(2, 360, 640), (1, 720, 1280), (2, 384, 384), (1, 200, 384), (2, 384, 500),
                                   (1, 97, 131), (1, 1080, 607)])
def test_resize_matches_pil_bit_exact(shape):
    rng = np.random.default_rng(sum(shape))
    frames = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    frames[0, :7, :9] = 255                                   # saturated / black patches: the clip8 path
    frames[0, -5:, -11:] = 0
    got = po.resize_bicubic_u8(frames, 384, 384)
    assert np.array_equal(got, _pil_resize(frames, 384, 384))


def test_full_preprocess_matches_transformers():
    from functools import partial, reduce
    from transformers.image_transforms import (convert_to_rgb, normalize, rescale, resize, to_channel_dimension_format)
    from transformers.image_utils import ChannelDimension, PILImageResampling, to_numpy_array
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(3, 300, 420, 3), dtype=np.uint8)
    fmt = ChannelDimension.FIRST
    transforms = [convert_to_rgb, to_numpy_array,                                  # siglip_encoder.py:55-62
                  partial(resize, size=(384, 384), resample=PILImageResampling.BICUBIC, data_format=fmt),
                  partial(rescale, scale=1 / 255, data_format=fmt),
                  partial(normalize, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), data_format=fmt),
                  partial(to_channel_dimension_format, channel_dim=fmt, input_channel_dim=fmt)]
    ref = np.stack(reduce(lambda x, f: [*map(f, x)], transforms, [to_numpy_array(f) for f in frames]))
    got = po.preprocess(frames)
    assert got.dtype == np.float32 and got.shape == (3, 3, 384, 384)
    assert np.array_equal(got, ref)
