"""Prediction-path host logic (CPU): the reference's own tests for `_blend_in` (packages/viscy-utils/tests/
test_prediction_writer.py) and `AugmentedPredictionVSUNet` (applications/cytoland/tests/test_engine.py:97-250) restated
against viscy_b200.predict, plus DivisiblePad / centre-crop round trips."""
import numpy as np
import pytest
import torch

from viscy_b200 import FullyConvolutionalMAE, Unet3d, UNeXt2
from viscy_b200.predict import (AugmentedPredictionVSUNet, DivisiblePad, _blend_in, _center_crop_to_shape,
                                _make_divisible_pad, rotation_tta_transforms)


def test_blend_in_consistency():
    depth = 5
    rng = np.random.default_rng(42)
    old_np = rng.random((2, depth, 8, 8)).astype(np.float32)
    new_np = rng.random((2, depth, 8, 8)).astype(np.float32)
    z_slice = slice(2, 2 + depth)
    result_np = _blend_in(old_np, new_np, z_slice)
    result_torch = _blend_in(torch.from_numpy(old_np).unsqueeze(0), torch.from_numpy(new_np).unsqueeze(0), z_slice)
    np.testing.assert_allclose(result_np, result_torch.squeeze(0).numpy(), rtol=1e-5, atol=1e-5)
    # known answer: window at z=2, depth 5 -> samples 3, factors [3, 3, 3, 2, 1]
    f = np.array([3, 3, 3, 2, 1], dtype=np.float32)[None, :, None, None]
    np.testing.assert_allclose(result_np, old_np * (f - 1) / f + new_np / f, rtol=1e-6)


def test_blend_in_zero_start_and_dtype():
    old, new = np.ones((2, 5, 8, 8), dtype=np.float32), np.zeros((2, 5, 8, 8), dtype=np.float32)
    np.testing.assert_array_equal(_blend_in(old, new, slice(0, 5)), new)
    r = _blend_in(torch.ones(1, 2, 5, 8, 8), torch.zeros(1, 2, 5, 8, 8), slice(2, 7))
    assert isinstance(r, torch.Tensor) and r.dtype == torch.float32


def test_divisible_pad_and_crop_round_trip():
    m = torch.nn.Identity()
    m.num_blocks = 3
    pad = _make_divisible_pad(m)
    x = torch.randn(2, 1, 5, 30, 45)
    p = pad(x)
    assert p.shape == (2, 1, 5, 32, 48)
    assert torch.equal(_center_crop_to_shape(p, x.shape[2:]), x)
    assert pad.widths(x.shape) == [(0, 0), (0, 0), (1, 1), (1, 2)]  # symmetric, the odd voxel goes behind
    m.downsamples_z = True
    assert _make_divisible_pad(m)(x).shape == (2, 1, 8, 32, 48)
    assert DivisiblePad((0, 0, 8, 8))(torch.zeros(1, 1, 1, 16, 16)).shape == (1, 1, 1, 16, 16)
    with pytest.raises(ValueError, match="Cannot crop dimension"):
        _center_crop_to_shape(x, (5, 31, 45))


def test_rotation_tta_transforms():
    forward, inverse = rotation_tta_transforms()
    assert len(forward) == len(inverse) == 4
    x = torch.randn(1, 1, 5, 6, 8)
    for fwd_t, inv_t in zip(forward, inverse):
        assert torch.allclose(inv_t(fwd_t(x)), x)
    with pytest.raises(ValueError, match="n must be >= 1"):
        rotation_tta_transforms(0)


def test_fnet3d_predict_sliding_windows():
    model = Unet3d(in_channels=1, out_channels=1, depth=1, mult_chan=8, in_stack_depth=4)
    vs = AugmentedPredictionVSUNet(model=model).eval()
    with torch.inference_mode():
        out = vs.predict_sliding_windows(torch.randn(1, 1, 8, 16, 16), out_channel=1, step=1)
    assert out.shape == (1, 1, 8, 16, 16)


def _fcmae(out_channels, z_window, blocks=2):
    return FullyConvolutionalMAE(in_channels=1, out_channels=out_channels, encoder_blocks=[2, 2, 2, 2], dims=[4, 8, 16, 32],
                                 decoder_conv_blocks=blocks, stem_kernel_size=[z_window, 4, 4], in_stack_depth=z_window,
                                 pretraining=False)


def test_predict_sliding_windows_output_and_blend_semantics():
    """Shape (test_engine.py:163-188) and values: the windows blended by hand with _blend_in."""
    torch.manual_seed(0)
    vs = AugmentedPredictionVSUNet(model=_fcmae(2, 5)).eval()
    x = torch.randn(1, 1, 12, 64, 64)
    with torch.inference_mode():
        out = vs.predict_sliding_windows(x, out_channel=2, step=1)
        ref = x.new_zeros(1, 2, 12, 64, 64)
        for s in range(0, 8):
            z = slice(s, s + 5)
            ref[:, :, z] = _blend_in(ref[:, :, z], vs.predict_step({"source": x[:, :, z]}), z)
    assert out.shape == (1, 2, 12, 64, 64)
    torch.testing.assert_close(out, ref)


def test_predict_sliding_windows_errors():
    vs = AugmentedPredictionVSUNet(model=UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto"))
    with pytest.raises(ValueError, match="5 dimensions"):
        vs.predict_sliding_windows(torch.randn(1, 5, 64, 64))
    with pytest.raises(ValueError, match="in_stack_depth 5 > input depth 3"):
        vs.predict_sliding_windows(torch.randn(1, 1, 3, 64, 64))
    lin = torch.nn.Linear(10, 10)
    lin.num_blocks = 1
    with pytest.raises(ValueError, match="out_stack_depth"):
        AugmentedPredictionVSUNet(model=lin).predict_sliding_windows(torch.randn(1, 1, 10, 4, 4))
    with pytest.raises(NotImplementedError, match="Only the 'predict' stage"):
        vs.setup("fit")


@pytest.mark.parametrize("yx", [(64, 64), (64, 48), (48, 64)])
def test_predict_sliding_windows_rotation_tta_nonsquare(yx):
    """test_engine.py:223-250: rotation TTA + sliding windows on non-square fields of view."""
    torch.manual_seed(0)
    vs = AugmentedPredictionVSUNet.with_rotation_tta(_fcmae(2, 5, blocks=1), reduction="median").eval()
    x = torch.randn(1, 1, 8, *yx)
    with torch.inference_mode():
        out = vs.predict_sliding_windows(x, out_channel=2, step=1)
    assert out.shape == (1, 2, 8, *yx) and torch.isfinite(out).all()
