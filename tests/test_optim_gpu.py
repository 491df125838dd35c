"""viscy_b200.optim.AdamW against torch.optim.AdamW (the optimizer the reference configures: VU/optimizers.py:10-61,
CY/engine.py:547-554): same updates over several steps on a real model's parameter set, unused parameters skipped,
state_dict interchange, GradScaler protocol (unscale + skip on inf), CUDA-graph capture."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _models():
    from viscy_b200 import UNeXt2
    torch.manual_seed(0)
    a = UNeXt2(in_channels=1, out_channels=2, in_stack_depth=5, backbone="convnextv2_atto").cuda()
    return a, copy.deepcopy(a)


def _fill_grads(ma, mb, seed, scale=1.0, skip=()):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for (n, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        if any(s in n for s in skip):
            pa.grad = pb.grad = None
            continue
        gr = torch.randn(pa.shape, device="cuda", generator=g) * scale
        pa.grad, pb.grad = gr.clone(), gr.clone()


@pytest.mark.parametrize("wd", [0.01, 0.0])
def test_adamw_matches_torch(wd):
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    oa = AdamW(ma.parameters(), lr=2e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    ob = torch.optim.AdamW(mb.parameters(), lr=2e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    from viscy_b200 import _lib as L
    n0 = L.launch_count()
    for step in range(5):
        _fill_grads(ma, mb, 10 + step, scale=10.0 ** (step - 2), skip=("stem",) if step < 2 else ())
        oa.step()
        ob.step()
    assert L.launch_count() - n0 == 5  # one launch per step for the whole group
    worst = 0.0
    for (n, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-6, atol=1e-7, msg=lambda m: f"{n}: {m}")
        sa, sb = oa.state[pa], ob.state[pb]
        for k in ("exp_avg", "exp_avg_sq"):  # torch's foreach path rounds its lerp differently: compare at the tensor's scale
            assert float((sa[k] - sb[k]).abs().max()) <= 2e-6 * float(sb[k].abs().max()), (n, k)
        worst = max(worst, float((pa.detach() - pb.detach()).abs().max()))
    print(f"\nAdamW wd={wd}: worst |dp| vs torch after 5 steps {worst:.2e}")
    steps = {n: float(oa.state[p]["step"]) for n, p in ma.named_parameters()}
    assert all(v == (3.0 if "stem" in n else 5.0) for n, v in steps.items())  # late joiners count their own steps


def test_state_dict_interchange():
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    ob = torch.optim.AdamW(mb.parameters(), lr=1e-3)
    for step in range(2):
        _fill_grads(ma, mb, 20 + step)
        for p in ma.parameters():
            p.grad = None
        ob.step()
    ma.load_state_dict(mb.state_dict())
    oa = AdamW(ma.parameters(), lr=1e-3)
    oa.load_state_dict(copy.deepcopy(ob.state_dict()))
    _fill_grads(ma, mb, 30)
    oa.step()
    ob.step()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-6, atol=1e-7)
    sd = oa.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 3.0
    oc = torch.optim.AdamW(mb.parameters(), lr=1e-3)
    oc.load_state_dict(sd)  # and back into torch's optimizer


def test_grad_scaler_protocol():
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    oa = AdamW(ma.parameters(), lr=1e-3)
    ob = torch.optim.AdamW(mb.parameters(), lr=1e-3, fused=True)
    sa = torch.amp.GradScaler("cuda", init_scale=1024.0)
    sb = torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(3):
        _fill_grads(ma, mb, 40 + step, scale=1024.0)
        if step == 1:  # an overflow: both must skip the step and halve the scale
            next(ma.parameters()).grad.view(-1)[0] = float("inf")
            next(mb.parameters()).grad.view(-1)[0] = float("inf")
        for sc, opt in ((sa, oa), (sb, ob)):
            if step == 0:
                sc._lazy_init_scale_growth_tracker(torch.device("cuda"))
            sc.step(opt)
            sc.update()
    assert sa.get_scale() == sb.get_scale() == 512.0
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-6, atol=1e-7)
    assert float(oa.state[next(ma.parameters())]["step"]) == 2.0


def test_capturable():
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    oa = AdamW(ma.parameters(), lr=1e-3)
    ob = torch.optim.AdamW(mb.parameters(), lr=1e-3)
    _fill_grads(ma, mb, 50)
    static = [p.grad for p in ma.parameters()]
    oa.step()
    ob.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        oa.step()
    ob.step()  # the capture itself does not run; the first replay is step 2
    for _ in range(2):
        graph.replay()
    ob.step()
    del static
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=5e-6, atol=2e-7)
    assert float(oa.state[next(ma.parameters())]["step"]) == 3.0


def test_loud_failures():
    from viscy_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError):
        AdamW([p]).step()
    q = torch.nn.Parameter(torch.zeros(8, device="cuda", dtype=torch.bfloat16))
    q.grad = torch.ones_like(q)
    with pytest.raises(NotImplementedError):
        AdamW([q]).step()


def test_param_groups_tensor_lr_and_maximize():
    """Two parameter groups with their own hyper-parameters, a device-tensor learning rate (what a capturable scheduler
    writes into) and maximize=True, against torch.optim.AdamW."""
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    pa, pb = list(ma.parameters()), list(mb.parameters())
    lr_t = torch.tensor(3e-3, device="cuda")
    oa = AdamW([{"params": pa[:20], "lr": lr_t, "weight_decay": 0.0}, {"params": pa[20:], "betas": (0.8, 0.95)}], lr=1e-3, maximize=True)
    ob = torch.optim.AdamW([{"params": pb[:20], "lr": 3e-3, "weight_decay": 0.0}, {"params": pb[20:], "betas": (0.8, 0.95)}], lr=1e-3,
                           maximize=True)
    for step in range(3):
        _fill_grads(ma, mb, 60 + step)
        if step == 2:  # the scheduler's move
            lr_t.fill_(1e-3)
            ob.param_groups[0]["lr"] = 1e-3
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7)


def test_capture_before_first_step_is_refused():
    """State and launch tables are created by an eager step; creating them under capture would replay the zero-fills."""
    from viscy_b200.optim import AdamW
    ma, mb = _models()
    oa = AdamW(ma.parameters(), lr=1e-3)
    _fill_grads(ma, mb, 70)
    graph = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="eager step"):
        with torch.cuda.graph(graph):
            oa.step()
    torch.cuda.synchronize()
    oa.step()  # and the optimizer is still usable eagerly
