"""Prediction path on the GPU: the fused crop + blend kernel vs the reference ops, and sliding-window inference (with
and without rotation TTA, padded non-divisible fields of view) through the sm_100a model vs the same wrapper on CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("odt,pdt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.float16, torch.float16)])
@pytest.mark.parametrize("start", [0, 1, 3, 7])
def test_blend_window_kernel_vs_reference_ops(cuda, odt, pdt, start):
    from viscy_b200.predict import _blend_in, _center_crop_to_shape, blend_window_
    g = torch.Generator(device=cuda).manual_seed(start)
    out = torch.rand(2, 3, 12, 30, 45, device=cuda, generator=g).to(odt)
    pred = torch.rand(2, 3, 5, 32, 48, device=cuda, generator=g).to(pdt)
    z = slice(start, start + 5)
    ref = out.float().clone()
    ref[:, :, z] = _blend_in(ref[:, :, z], _center_crop_to_shape(pred.float(), (30, 45)), z)
    blend_window_(out, pred, start)
    tol = 1e-6 if odt == torch.float32 else 1e-3
    assert rel(out.float(), ref) < tol
    untouched = torch.ones(12, dtype=torch.bool)
    untouched[z] = False
    assert torch.equal(out[:, :, untouched].float(), ref[:, :, untouched].to(odt).float())


@pytest.mark.parametrize("tta", [False, True])
def test_sliding_windows_unext2_vs_cpu(cuda, tta):
    from viscy_b200 import UNeXt2
    from viscy_b200.predict import AugmentedPredictionVSUNet
    torch.manual_seed(0)
    cfg = dict(in_channels=1, out_channels=2, in_stack_depth=5, backbone="convnextv2_atto", stem_kernel_size=(5, 4, 4))
    m = UNeXt2(**cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)
    make = (lambda mod: AugmentedPredictionVSUNet.with_rotation_tta(mod, reduction="mean")) if tta else AugmentedPredictionVSUNet
    x = torch.randn(1, 1, 8, 80, 112)  # neither 80 nor 112 is divisible by 2**6: padded to 128 x 128, cropped back
    with torch.inference_mode():
        ref = make(m).eval().predict_sliding_windows(x, out_channel=2, step=1)
    mg = UNeXt2(**cfg).eval()
    mg.load_state_dict(m.state_dict())
    mg = mg.to(cuda)
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.float16):
        out = make(mg).eval().predict_sliding_windows(x.to(cuda), out_channel=2, step=1)
    e = rel(out.float().cpu(), ref)
    print(f"\nsliding windows (tta={tta}) rel-L2 vs CPU {e:.3e}")
    assert out.shape == (1, 2, 8, 80, 112) and out.dtype == torch.float32
    assert e < 3e-3
