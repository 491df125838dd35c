"""MixedLoss / ms_ssim_25d through the fused sm_100a kernels vs the reference-generated goldens (100 % reference code) and,
element-wise for the prediction gradient, vs the fp32 CPU mirror (== the reference, tests/test_mixed_loss_cpu.py).
Tolerances: the reference rounds the five window means to bf16, so its own SSIM maps carry ~1e-2 per-pixel noise; a
different (fp32) summation order flips a small fraction of those roundings.  The reference's own test bounds the autocast
drift of the scalar by rtol = atol = 1e-2 (test_mixed_loss.py:45-68); the scalars here are held to 2e-3."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", ["mixed_loss_default", "mixed_loss_all"])
def test_mixed_loss_against_reference_golden(cuda, name):
    from viscy_b200.losses import MixedLoss, ms_ssim_25d, ssim_25d
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    x = g["x"].float().to(cuda).requires_grad_(True)
    y = g["y"].float().to(cuda)
    loss = MixedLoss(**g["kw"])(x, y)
    loss.backward()
    print(f"\n[{name}] loss {loss.item():.6f} vs {g['loss']:.6f}; grad norm {x.grad.norm().item():.4e} vs {g['grad_norm']:.4e}")
    assert loss.dtype == torch.float32 and loss.ndim == 0
    assert abs(loss.item() - g["loss"]) < 2e-3 * abs(g["loss"])
    assert abs(ms_ssim_25d(x.detach(), y, clamp=True).item() - g["ms_ssim"]) < 2e-3
    s, c = ssim_25d(x.detach(), y, return_contrast_sensitivity=True)
    torch.testing.assert_close(s.cpu(), g["ssim"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(c.cpu(), g["cs"], rtol=2e-3, atol=2e-3)
    # gradient: the reference's backward runs its convolutions in bf16; the straight-through fp32 gradient of the same
    # (rounded) forward is the yardstick: ours sits on it, the reference's scatters around it
    from ssim_straight_through import mixed
    xi = g["x"].float().requires_grad_(True)
    mixed(xi, g["y"].float(), g["kw"]).backward()
    ideal = xi.grad[..., ::3, ::3]
    e_ref = rel(g["grad_sub"], ideal)
    e_ours = rel(x.grad[..., ::3, ::3].cpu(), ideal)
    e = rel(x.grad[..., ::3, ::3].cpu(), g["grad_sub"])
    print(f"gradient rel-L2: ours vs straight-through fp32 {e_ours:.3e}, reference vs the same {e_ref:.3e}, ours vs reference {e:.3e}")
    assert e_ours < 1e-2 and e < 1.25 * e_ref + 1e-3
    assert abs(x.grad.norm().item() - g["grad_norm"]) < 2e-2 * g["grad_norm"]


def _emulated_autocast_loss(pred16, target, kw):
    """The reference math on CPU with the prediction held in `pred16.dtype` between pyramid levels (what CUDA autocast
    does: avg_pool3d keeps the input dtype); CPU torch has no 16-bit avg_pool3d, so the pooling runs in fp32 and is
    rounded."""
    from viscy_b200.losses import _combine, ssim_25d
    loss = 0
    if kw["l1_alpha"]:
        loss = loss + F.l1_loss(pred16.float(), target) * kw["l1_alpha"]
    if kw["l2_alpha"]:
        loss = loss + F.mse_loss(pred16.float(), target) * kw["l2_alpha"]
    cs_list, p, t = [], pred16, target
    for _ in range(5):
        ssim, cs = ssim_25d(p.float(), t, return_contrast_sensitivity=True)
        cs_list.append(cs)
        p = F.avg_pool3d(p.float(), (1, 2, 2)).to(pred16.dtype)
        t = F.avg_pool3d(t, (1, 2, 2))
    ms = _combine(ssim, cs_list, True, (0.0448, 0.2856, 0.3001, 0.2363, 0.1333))
    return loss + (1 - ms) * kw["ms_dssim_alpha"]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_mixed_loss_16bit_prediction(cuda, dtype):
    """bf16 / fp16 predictions (the model output under autocast) with an fp32 target, odd plane sizes, two channels."""
    from viscy_b200.losses import MixedLoss
    kw = dict(l1_alpha=0.5, l2_alpha=0.25, ms_dssim_alpha=0.5)
    g = torch.Generator().manual_seed(5)
    p = torch.rand(2, 2, 3, 182, 210, generator=g)
    t = 0.5 * p + 0.5 * torch.rand(2, 2, 3, 182, 210, generator=g)
    pc = p.to(dtype).requires_grad_(True)
    ref = _emulated_autocast_loss(pc, t, kw)
    ref.backward()
    x = p.to(dtype).to(cuda).requires_grad_(True)
    with torch.autocast("cuda", dtype=dtype):
        loss = MixedLoss(**kw)(x, t.to(cuda))
    loss.backward()
    e = rel(x.grad.float().cpu(), pc.grad.float())
    print(f"\n[{dtype}] loss {loss.item():.6f} vs {ref.item():.6f}; gradient rel-L2 {e:.3e}")
    assert loss.dtype == torch.float32 and x.grad.dtype == dtype
    assert abs(loss.item() - ref.item()) < 2e-3 * abs(ref.item())
    assert e < 8e-2  # the CPU reference gradient itself carries 4-6e-2 of bf16 noise (see the golden test)


def test_l1_l2_only_and_reference_test_cases(cuda):
    """packages/viscy-utils/tests/test_mixed_loss.py restated: finite fp32 scalar in / outside autocast, autocast drift
    within rtol = atol = 1e-2, finite gradients, L1-only == alpha * F.l1_loss."""
    from viscy_b200.losses import MixedLoss
    torch.manual_seed(0)
    pred, target = torch.rand(2, 1, 15, 192, 192, device=cuda), torch.rand(2, 1, 15, 192, 192, device=cuda)
    fn = MixedLoss(l1_alpha=0.5, l2_alpha=0.0, ms_dssim_alpha=0.5)
    loss = fn(pred, target)
    assert loss.dtype == torch.float32 and loss.ndim == 0 and torch.isfinite(loss).item()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la = fn(pred, target)
    torch.testing.assert_close(la, loss, rtol=1e-2, atol=1e-2)
    pr = pred.detach().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fn(pr, target).backward()
    assert pr.grad is not None and pr.grad.shape == pr.shape and torch.isfinite(pr.grad).all().item()
    p2, t2 = torch.rand(2, 1, 8, 64, 64, device=cuda), torch.rand(2, 1, 8, 64, 64, device=cuda)
    p2.requires_grad_(True)
    l1 = MixedLoss(0.5, 0.0, 0.0)(p2, t2)
    torch.testing.assert_close(l1, F.l1_loss(p2, t2) * 0.5, rtol=1e-6, atol=0)
    l1.backward()
    torch.testing.assert_close(p2.grad, 0.5 * torch.sign(p2.detach() - t2) / p2.numel(), rtol=1e-6, atol=0)
    l12 = MixedLoss(0.25, 0.75, 0.0)(p2.detach(), t2)
    torch.testing.assert_close(l12, 0.25 * F.l1_loss(p2, t2) + 0.75 * F.mse_loss(p2, t2), rtol=1e-5, atol=0)


def test_window_too_large_raises(cuda):
    from viscy_b200.losses import ms_ssim_25d
    with pytest.raises(RuntimeError, match="smaller than the 11x11 SSIM window"):
        ms_ssim_25d(torch.rand(1, 1, 2, 64, 64, device=cuda), torch.rand(1, 1, 2, 64, 64, device=cuda))


def test_mixed_loss_step_is_cuda_graph_capturable(cuda):
    """Forward + backward of the loss inside one CUDA graph (no host copies or synchronisation on the path): replays give
    the eager value and gradient."""
    import warnings
    from viscy_b200.graphs import GraphedStep
    from viscy_b200.losses import MixedLoss
    warnings.filterwarnings("ignore", message="Input depth")
    g = torch.Generator(device=cuda).manual_seed(9)
    x = torch.rand((2, 2, 5, 192, 224), device=cuda, generator=g).bfloat16()
    y = torch.rand((2, 2, 5, 192, 224), device=cuda, generator=g)
    crit = MixedLoss()
    grads = []

    def step(xd, yd):
        xd = xd.detach().requires_grad_(True)
        loss = crit(xd, yd)
        loss.backward()
        grads.append(xd.grad)
        return loss

    eager = step(x, y)
    g_eager = grads[-1].clone()
    gs = GraphedStep(step, (x, y), warmup=2)
    out = gs(x, y)
    torch.cuda.synchronize()
    assert abs(out.item() - eager.item()) < 1e-4 * abs(eager.item())  # fp32 atomic sums: order differs run to run
    assert rel(grads[-1].float(), g_eager.float()) < 2e-3  # bf16 gradient: last-bit differences from the atomic sum order


@pytest.mark.parametrize("reduction", ["mean", "sum"])
@pytest.mark.parametrize("pdt,tdt,shape", [(torch.bfloat16, torch.float32, (2, 2, 5, 64, 64)), (torch.float16, torch.float16, (3, 1, 7, 33)),
                                           (torch.float32, torch.float32, (1, 1003)), (torch.bfloat16, torch.bfloat16, (5,)),
                                           (torch.float16, torch.float32, (8, 2, 21, 64, 64))])
def test_mse_loss_matches_torch(pdt, tdt, shape, reduction):
    """viscy_b200.losses.MSELoss == torch mse_loss on the fp32 copies (what autocast computes): scalar and gradient, ragged
    sizes (tails of the 8-element vectors), every dtype pair, under a scaled upstream gradient."""
    from viscy_b200.losses import MSELoss
    g = torch.Generator(device="cuda").manual_seed(11)
    p = torch.randn(shape, device="cuda", generator=g).to(pdt).requires_grad_(True)
    t = torch.randn(shape, device="cuda", generator=g).to(tdt)
    loss = MSELoss(reduction)(p, t)
    (loss * 3.0).backward()
    pr = p.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(pr.float(), t.float(), reduction=reduction)
    (ref * 3.0).backward()
    assert loss.dtype == torch.float32 and loss.shape == ()
    torch.testing.assert_close(loss, ref, rtol=2e-5, atol=1e-7)
    assert p.grad.dtype == pdt
    # (fp16 gradients of a mean over ~1e6 elements are subnormal: allow two quanta of 2^-24)
    torch.testing.assert_close(p.grad.float(), pr.grad.float(), rtol=1e-2 if pdt != torch.float32 else 1e-6,
                               atol=1.2e-7 if pdt == torch.float16 else 1e-9)
