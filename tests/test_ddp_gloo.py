"""N > 1 host path on CPU: world_size-2 gloo DDP over the product module (torch backend) - gradients are the
all-reduced mean of the per-rank gradients, exactly as the reference's Lightning DDP strategy would produce."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, init_file, out_file):
    from viscy_b200 import UNeXt2
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto")
    ddp = torch.nn.parallel.DistributedDataParallel(m)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn((1, 1, 5, 32, 32), generator=g)
    y = torch.randn((1, 1, 5, 32, 32), generator=g)
    torch.nn.functional.mse_loss(ddp(x), y).backward()
    flat = torch.cat([p.grad.flatten() for p in m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"ranks_equal": bool(torch.equal(gathered[0], gathered[1])), "grad": flat}, out_file)
    dist.destroy_process_group()


def test_ddp_gloo_two_ranks():
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "init"), os.path.join(d, "out.pt")
        mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    assert res["ranks_equal"]
    # single-process reference: mean of the two per-rank gradients
    from viscy_b200 import UNeXt2
    grads = []
    for rank in range(2):
        torch.manual_seed(0)
        m = UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto")
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn((1, 1, 5, 32, 32), generator=g)
        y = torch.randn((1, 1, 5, 32, 32), generator=g)
        torch.nn.functional.mse_loss(m(x), y).backward()
        grads.append(torch.cat([p.grad.flatten() for p in m.parameters()]))
    torch.testing.assert_close(res["grad"], (grads[0] + grads[1]) / 2, rtol=1e-5, atol=1e-7)


def _flat_worker(rank, world, init_file, out_file):
    from viscy_b200 import UNeXt2
    from viscy_b200.parallel import FlatGradAllReduce
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(rank)  # different initial weights per rank: broadcast must equalise them
    m = UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto")
    ex = FlatGradAllReduce(m.parameters())
    ex.broadcast_parameters(0)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn((1, 1, 5, 32, 32), generator=g)
    y = torch.randn((1, 1, 5, 32, 32), generator=g)
    torch.nn.functional.mse_loss(m(x), y).backward()
    ex()
    if rank == 1:
        torch.save({"grad": torch.cat([p.grad.flatten() for p in m.parameters()]),
                    "w": torch.cat([p.detach().flatten() for p in m.parameters()])}, out_file)
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_ddp_semantics():
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "init"), os.path.join(d, "out.pt")
        mp.spawn(_flat_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    from viscy_b200 import UNeXt2
    grads = []
    for rank in range(2):
        torch.manual_seed(0)  # rank 0's weights everywhere after the broadcast
        m = UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto")
        if rank == 0:
            torch.testing.assert_close(res["w"], torch.cat([p.detach().flatten() for p in m.parameters()]))
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn((1, 1, 5, 32, 32), generator=g)
        y = torch.randn((1, 1, 5, 32, 32), generator=g)
        torch.nn.functional.mse_loss(m(x), y).backward()
        grads.append(torch.cat([p.grad.flatten() for p in m.parameters()]))
    torch.testing.assert_close(res["grad"], (grads[0] + grads[1]) / 2, rtol=1e-5, atol=1e-7)


def _bucket_worker(rank, world, init_file, out_file):
    from viscy_b200 import Unet25d, UNeXt2
    from viscy_b200.parallel import BucketedGradAllReduce
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    res = {}
    for name, make, shape in (("unext2", lambda: UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto"),
                               (1, 1, 5, 32, 32)),
                              ("unet25d", lambda: Unet25d(num_filters=(4, 8), num_blocks=1, dropout=0.0), (1, 1, 5, 32, 32))):
        torch.manual_seed(rank)
        m = make()
        ex = BucketedGradAllReduce(m.parameters(), fractions=(0.2, 0.5, 0.9))
        ex.broadcast_parameters(0)
        steps = []
        for step in range(3):  # step 0 records the backward order and builds the buckets; 1 and 2 overlap
            g = torch.Generator().manual_seed(100 + rank + 10 * step)
            x = torch.randn(shape, generator=g)
            m.zero_grad(set_to_none=True)
            out = m(x)
            torch.nn.functional.mse_loss(out, torch.randn(out.shape, generator=g)).backward()
            ex.finish()
            steps.append(torch.cat([(p.grad if p.grad is not None else torch.full_like(p, float("nan"))).flatten()
                                    for p in m.parameters()]))
        res[name] = {"grads": steps, "n_buckets": len(ex.buckets),
                     "none": [n for n, p in m.named_parameters() if p.grad is None]}
    if rank == 1:
        torch.save(res, out_file)
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_overlap_and_unused_parameters():
    """Bucketed exchange: every step (the recording one and the overlapped ones) yields the mean of the per-rank
    gradients; parameters without a gradient (Unet25d's unused resid_conv) keep grad=None as under stock DDP."""
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "init"), os.path.join(d, "out.pt")
        mp.spawn(_bucket_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    from viscy_b200 import Unet25d, UNeXt2
    for name, make, shape in (("unext2", lambda: UNeXt2(in_channels=1, out_channels=1, in_stack_depth=5, backbone="convnextv2_atto"),
                               (1, 1, 5, 32, 32)),
                              ("unet25d", lambda: Unet25d(num_filters=(4, 8), num_blocks=1, dropout=0.0), (1, 1, 5, 32, 32))):
        assert res[name]["n_buckets"] >= 3
        for step in range(3):
            grads = []
            for rank in range(2):
                torch.manual_seed(0)
                m = make()
                g = torch.Generator().manual_seed(100 + rank + 10 * step)
                x = torch.randn(shape, generator=g)
                out = m(x)
                torch.nn.functional.mse_loss(out, torch.randn(out.shape, generator=g)).backward()
                grads.append(torch.cat([(p.grad if p.grad is not None else torch.full_like(p, float("nan"))).flatten()
                                        for p in m.parameters()]))
            torch.testing.assert_close(res[name]["grads"][step], (grads[0] + grads[1]) / 2, rtol=1e-5, atol=1e-7,
                                       equal_nan=True)
    assert any("resid_conv" in n for n in res["unet25d"]["none"]) and not res["unext2"]["none"]
