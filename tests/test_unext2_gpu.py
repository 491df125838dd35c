"""UNeXt2 through the sm_100a kernels vs the fp32 CPU oracle (oracle/models.py) on the same seeded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(in_channels=1, out_channels=2, in_stack_depth=14, backbone="convnextv2_tiny",
           stem_kernel_size=(7, 4, 4), head_pool=True)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _pair(cuda, cfg, seed=0):
    from oracle import models as OM
    from viscy_b200 import UNeXt2
    torch.manual_seed(seed)
    o = OM.UNeXt2(**cfg)
    # GRN weights are zero-initialised: give them (and biases) values so every path carries signal
    with torch.no_grad():
        for n, p in o.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)
    m = UNeXt2(**cfg)
    m.load_state_dict(o.state_dict())
    return o, m.to(cuda)


# absolute bounds; the per-tensor bound relative to stock autocast lives in test_parity_yardstick_gpu.py
@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float16, 1.5e-3, 3e-2), (torch.bfloat16, 1e-2, 8e-2)])
def test_unext2_fwd_bwd_parity(cuda, dtype, ftol, gtol):
    o, m = _pair(cuda, CFG)
    torch.manual_seed(1)
    x = torch.randn(2, 1, 14, 64, 64)
    tgt = torch.randn(2, 2, 14, 64, 64)
    ref = o(x)
    torch.nn.functional.mse_loss(ref, tgt).backward()
    with torch.autocast("cuda", dtype=dtype):
        out = m(x.to(cuda))
        loss = torch.nn.functional.mse_loss(out.float(), tgt.to(cuda))
    scale = 65536.0 if dtype == torch.float16 else 1.0  # static GradScaler stand-in
    (loss * scale).backward()
    for p in m.parameters():
        p.grad.div_(scale)
    assert out.shape == ref.shape and out.dtype == dtype
    e = rel(out.float().cpu(), ref)
    print(f"\n[{dtype}] forward rel-L2 {e:.3e}")
    assert e < ftol
    og = dict(o.named_parameters())
    worst = []
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert p.grad.shape == p.shape
        gref = og[n].grad
        if n == "head.conv.0.conv.bias":  # bias in front of InstanceNorm: analytically zero gradient
            assert p.grad.float().norm().item() < 1e-2 and gref.norm().item() < 1e-2, n
            continue
        worst.append((rel(p.grad.cpu(), gref), n))
    worst.sort(reverse=True)
    print("worst grads:", [(f"{w:.2e}", n) for w, n in worst[:6]])
    import statistics
    print("median grad rel-L2", statistics.median(w for w, _ in worst))
    assert worst[0][0] < gtol


def test_unext2_eval_no_grad_and_state_dict(cuda):
    o, m = _pair(cuda, CFG)
    assert list(o.state_dict()) == list(m.state_dict())
    m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        y = m(torch.randn(1, 1, 14, 64, 64, device=cuda))
    assert y.shape == (1, 2, 14, 64, 64) and torch.isfinite(y).all()


def test_unext2_requires_16bit(cuda):
    _, m = _pair(cuda, CFG)
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 1, 14, 64, 64, device=cuda))


def test_unext2_full_config_step(cuda):
    """BASELINE config 2 shape (B=8, 21x256x256 bf16): forward against the fp32 oracle moved to the GPU (TF32 off), loss
    against its loss, finite gradients, and a second step on the same buffers gives the same result (up to the order of fp32 atomic sums)."""
    from oracle import models as OM
    from viscy_b200 import UNeXt2
    cfg = dict(in_channels=1, out_channels=2, in_stack_depth=21, backbone="convnextv2_tiny",
               stem_kernel_size=(7, 4, 4), head_pool=True)
    torch.manual_seed(0)
    o = OM.UNeXt2(**cfg)
    m = UNeXt2(**cfg)
    m.load_state_dict(o.state_dict())
    m, o = m.to(cuda), o.to(cuda)
    x = torch.randn(8, 1, 21, 256, 256, device=cuda)
    tgt = torch.randn(8, 2, 21, 256, 256, device=cuda)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref = torch.cat([o(x[i:i + 2]) for i in range(0, 8, 2)])
    ref_loss = torch.nn.functional.mse_loss(ref, tgt).item()
    outs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(x)
            loss = torch.nn.functional.mse_loss(out.float(), tgt)
        loss.backward()
        outs.append(out.float())
    torch.cuda.synchronize()
    assert out.shape == (8, 2, 21, 256, 256)
    e = rel(out.float().cpu(), ref.cpu())
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):  # the yardstick: stock autocast on the same oracle
        stock = torch.cat([o(x[i:i + 2]) for i in range(0, 8, 2)]).float()
    es = rel(stock.cpu(), ref.cpu())
    print(f"\nfull-config forward rel-L2 vs fp32 oracle: {e:.3e} (stock autocast {es:.3e}); loss {loss.item():.6f} vs {ref_loss:.6f}")
    assert e <= 1.5 * es + 2e-4 and e < 1.2e-2
    assert abs(loss.item() - ref_loss) < 2e-3 * abs(ref_loss)
    assert rel(outs[0].cpu(), outs[1].cpu()) < 2e-3  # same buffers, same result (fp32 atomics reorder the GRN sums)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("name", ["unext2_atto", "unext2_tiny"])
def test_unext2_against_reference_golden(cuda, name):
    """sm_100a path vs the committed golden vectors produced by the reference's own code (tests/golden/)."""
    from pathlib import Path
    from oracle import models as OM
    from viscy_b200 import UNeXt2
    g = torch.load(Path(__file__).resolve().parent / "golden" / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    o = OM.UNeXt2(**g["cfg"])  # consumes the RNG exactly like the reference: same weights as the golden run
    m = UNeXt2(**g["cfg"])
    m.load_state_dict(o.state_dict())
    m = m.to(cuda)
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(g["x"].to(cuda))
        loss = torch.nn.functional.mse_loss(out.float(), g["targets"][0].to(cuda))
    loss.backward()
    e = rel(out.float().cpu(), g["outs"][0])
    print(f"\n[{name}] forward rel-L2 vs reference golden {e:.3e}")
    assert e < 2e-3  # 1e-3-class fp16 tolerance of north_star, L2 over the whole output
    assert abs(loss.item() - g["loss"]) < 2e-3 * abs(g["loss"])
    bad = []
    for n, p in m.named_parameters():
        ref = g["grad_norms"][n]
        if ref < 1e-6 or n == "head.conv.0.conv.bias":  # bias before InstanceNorm: analytically zero gradient
            continue
        if abs(p.grad.float().norm().item() - ref) > 5e-2 * ref:
            bad.append((n, p.grad.float().norm().item(), ref))
    assert not bad, bad[:5]
