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
