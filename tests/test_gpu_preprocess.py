"""GPU parity of the frame pre-processing kernels (csrc/preprocess.cu) against the oracle: bit-exact."""
import numpy as np
import pytest
import torch

import mavlm_b200 as M
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 360, 640), (1, 720, 1280), (2, 384, 384), (1, 200, 384), (2, 384, 500),
                                   (1, 97, 131), (0, 64, 64)])
def test_preprocess_bit_exact(shape):
    rng = np.random.default_rng(sum(shape))
    frames = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    if shape[0]:
        frames[0, :7, :9] = 255
        frames[0, -5:, -11:] = 0
    x, u8 = M.preprocess.frames_preprocess(torch.from_numpy(frames).cuda(), return_resized=True)
    torch.cuda.synchronize()
    assert np.array_equal(u8.cpu().numpy(), po.resize_bicubic_u8(frames, 384, 384))
    assert np.array_equal(x.cpu().numpy(), po.preprocess(frames))


def test_processor_interface_and_bf16():
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, size=(4, 240, 320, 3), dtype=np.uint8)
    proc = M.SigLipImageProcessor(dtype=torch.bfloat16)
    out = proc.preprocess(torch.from_numpy(frames), return_tensors="pt")["pixel_values"]       # CPU tensor in: H2D copy
    ref = torch.from_numpy(po.preprocess(frames)).to(torch.bfloat16)                          # siglip_encoder.py:585
    assert out.dtype == torch.bfloat16 and out.shape == (4, 3, 384, 384)
    assert torch.equal(out.cpu(), ref)
    lst = proc.preprocess([f for f in frames])["pixel_values"]                                   # list of frames
    assert torch.equal(lst, out)
    half = M.SigLipImageProcessor(dtype=torch.float16).preprocess(frames)["pixel_values"]        # fp16 tower (builder.py:27)
    assert half.dtype == torch.float16 and torch.equal(half.cpu(), torch.from_numpy(po.preprocess(frames)).half())
    with pytest.raises(ValueError):
        M.preprocess.frames_preprocess(torch.zeros(2, 8, 8, 3, device="cuda"))                 # not uint8


def test_full_resolution_round_trip_property():
    """Size-independent property at a full-size input: resizing a constant image gives that constant (the
    coefficients of every output pixel sum to 1 in fixed point up to rounding, then clip8)."""
    for v in (0, 1, 127, 254, 255):
        frames = torch.full((2, 1080, 1920, 3), v, dtype=torch.uint8, device="cuda")
        _, u8 = M.preprocess.frames_preprocess(frames, return_resized=True)
        assert int(u8.min()) == v and int(u8.max()) == v
