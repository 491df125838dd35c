"""Host side of viscy_b200.optim.AdamW: constructor checks mirror torch.optim.AdamW, the launch tables cover every element
exactly once, and CPU parameters fail loudly (there is no CPU path)."""
import pytest
import torch


def test_constructor_checks():
    from viscy_b200.optim import AdamW
    p = [torch.nn.Parameter(torch.zeros(4))]
    for kw in (dict(lr=-1.0), dict(eps=-1e-8), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0)), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            AdamW(p, **kw)
    opt = AdamW(p, lr=3e-4, betas=(0.8, 0.9), eps=1e-6, weight_decay=0.0)
    g = opt.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"], g["maximize"]) == (3e-4, (0.8, 0.9), 1e-6, 0.0, False)
    assert AdamW._step_supports_amp_scaling


def test_cpu_parameters_fail_loudly():
    from viscy_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        AdamW([p]).step()
    AdamW([torch.nn.Parameter(torch.zeros(4))]).step()  # nothing to do without gradients


def test_chunk_table_covers_every_element_once():
    from viscy_b200 import optim
    sizes = [1, 7, 2048, 2049, 5000, 3]
    params = [torch.zeros(n) for n in sizes]
    plan_chunks, start = [], [0]
    for i, p in enumerate(params):  # the same walk _Plan does (it needs device tensors for the tables themselves)
        for off in range(0, p.numel(), optim.CHUNK):
            plan_chunks.append((i % optim.WINDOW, off, min(optim.CHUNK, p.numel() - off)))
        start.append(len(plan_chunks))
    for i, n in enumerate(sizes):
        cs = plan_chunks[start[i]:start[i + 1]]
        assert [c[1] for c in cs] == list(range(0, n, optim.CHUNK)) and sum(c[2] for c in cs) == n
        assert all(c[0] == i and 0 < c[2] <= optim.CHUNK for c in cs)


def test_flat_gradient_slots_are_16_byte_aligned():
    """parallel._Bucket pads every gradient's slot in the flat buffer to 4 elements, so the views handed back as `.grad`
    keep the one-launch AdamW on its float4 path even behind tensors with odd element counts (a 2-element bias)."""
    from viscy_b200.parallel import _Bucket, _slot
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((2,), (3, 5), (8,), (1,), (7, 7))]
    flat = torch.zeros(sum(_slot(p) for p in params))
    b = _Bucket(params, flat)
    assert [_slot(p) for p in params] == [4, 16, 8, 4, 52]
    for p, v in zip(params, b.views):
        assert v.shape == p.shape and (v.data_ptr() - flat.data_ptr()) % 16 == 0
    for v in b.views:
        v.fill_(1.0)
    assert float(flat.sum()) == sum(p.numel() for p in params)  # the pads stay zero
