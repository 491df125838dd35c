"""A LightningModule-shaped caller driving the sm_100a modules the way the reference's engines do (no Lightning in
this image, so the step is restated): dict batch -> model(source) -> loss -> backward under autocast, the model wrapped
in stock DistributedDataParallel over NCCL (world size 1), AdamW step.

Mirrors `VSUNet.training_step` (applications/cytoland/src/cytoland/engine.py:265-304: list of Sample dicts, `pred =
self.forward(source)`, `loss_function(pred, target)`, mean over the list) and `ContrastiveModule.training_step`
(applications/dynaclr/src/dynaclr/engine.py:262-287: two forwards, NT-Xent over the stacked projections).
The same step on the fp32 oracle (CPU) gives the reference loss curve; three optimizer steps must track it."""
import os

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


class _VSUNetLike(torch.nn.Module):
    """The part of VSUNet that touches the model: registry lookup -> net_class(**model_config), forward, MSE loss."""

    def __init__(self, registry, architecture, model_config):
        super().__init__()
        self.model = registry[architecture](**model_config)
        self.loss_function = torch.nn.MSELoss()

    def forward(self, x):
        return self.model(x)

    def training_step(self, batch, batch_idx):
        if not isinstance(batch, (list, tuple)):
            batch = [batch]
        losses = []
        for b in batch:
            pred = self.forward(b["source"])
            losses.append(self.loss_function(pred.float(), b["target"]))
        return torch.stack(losses).mean()


@pytest.fixture(scope="module")
def nccl_world1(cuda):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda)
    yield
    dist.destroy_process_group()


def test_vsunet_like_training_steps_under_ddp(cuda, nccl_world1):
    from oracle import models as OM
    import viscy_b200
    cfg = dict(in_channels=1, out_channels=2, in_stack_depth=14, backbone="convnextv2_tiny",
               stem_kernel_size=(7, 4, 4), head_pool=True)
    registry = {"UNeXt2": viscy_b200.UNeXt2, "2.5D": viscy_b200.Unet25d, "FNet3D": viscy_b200.Unet3d}
    torch.manual_seed(0)
    ref = _VSUNetLike({"UNeXt2": OM.UNeXt2}, "UNeXt2", cfg)
    mod = _VSUNetLike(registry, "UNeXt2", cfg)
    mod.load_state_dict(ref.state_dict())  # Lightning checkpoints carry the "model." prefix: same keys here
    assert all(k.startswith("model.") for k in mod.state_dict())
    mod = mod.to(cuda)
    ddp = torch.nn.parallel.DistributedDataParallel(mod, device_ids=[cuda.index], static_graph=True)
    opt = torch.optim.AdamW(mod.parameters(), lr=2e-4)
    ropt = torch.optim.AdamW(ref.parameters(), lr=2e-4)
    g = torch.Generator().manual_seed(5)
    losses, rlosses = [], []
    for step in range(3):
        batch = [{"source": torch.randn(1, 1, 14, 64, 64, generator=g), "target": torch.randn(1, 2, 14, 64, 64, generator=g)}
                 for _ in range(2)]  # a list of Sample dicts, as the reference's concatenated data modules produce
        ropt.zero_grad(set_to_none=True)
        rl = ref.training_step(batch, step)
        rl.backward()
        ropt.step()
        rlosses.append(rl.item())
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            # training_step with self.forward routed through the DDP wrapper (what Lightning's strategy does)
            loss = torch.stack([mod.loss_function(ddp(b["source"].to(cuda)).float(), b["target"].to(cuda)) for b in batch]).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("\nloss (sm_100a, bf16, DDP):", losses, "\nloss (fp32 oracle):       ", rlosses)
    for a, b in zip(losses, rlosses):
        assert abs(a - b) < 1e-2 * abs(b)
    # parameters after three AdamW steps still track the oracle's
    rp = dict(ref.named_parameters())
    worst = max((p.detach().cpu() - rp[n].detach()).abs().max().item() for n, p in mod.named_parameters())
    assert worst < 2e-3  # three steps of lr 2e-4: updates are <= 6e-4 per element


def test_contrastive_module_like_training_step(cuda, nccl_world1):
    import viscy_b200
    from oracle import models as OM
    from viscy_b200.loss import NTXentLoss
    cfg = dict(backbone="convnext_tiny", in_channels=2, in_stack_depth=15, stem_kernel_size=(5, 4, 4),
               stem_stride=(5, 4, 4), embedding_dim=768, projection_dim=128)
    torch.manual_seed(1)
    ref = OM.ContrastiveEncoder(**cfg)
    enc = viscy_b200.ContrastiveEncoder(**cfg)
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(cuda)
    ddp = torch.nn.parallel.DistributedDataParallel(enc, device_ids=[cuda.index], static_graph=True)
    crit = NTXentLoss(temperature=0.2)
    g = torch.Generator().manual_seed(9)
    batch = {"anchor": torch.randn(4, 2, 15, 64, 64, generator=g), "positive": torch.randn(4, 2, 15, 64, 64, generator=g)}
    labels = torch.cat([torch.arange(4), torch.arange(4)])

    def step(model, dev, dtype):
        with torch.autocast(dev.type, dtype=dtype, enabled=dtype is not None):
            _, pa = model(batch["anchor"].to(dev))
            _, pp = model(batch["positive"].to(dev))
        loss = crit(torch.cat([pa, pp]).float(), labels.to(dev))
        loss.backward()
        return loss.item()

    rl = step(ref, torch.device("cpu"), None)
    ml = step(ddp, cuda, torch.bfloat16)
    print(f"\nNT-Xent loss: sm_100a bf16 under DDP {ml:.5f}, fp32 oracle {rl:.5f}")
    assert abs(ml - rl) < 3e-2 * abs(rl)
    rg = dict(ref.named_parameters())
    big = [(n, p) for n, p in enc.named_parameters() if rg[n].grad is not None and rg[n].grad.norm() > 1e-4]
    cos = [torch.nn.functional.cosine_similarity(p.grad.flatten().cpu().float(), rg[n].grad.flatten(), dim=0).item() for n, p in big]
    assert min(cos) > 0.9, sorted(zip(cos, [n for n, _ in big]))[:5]
