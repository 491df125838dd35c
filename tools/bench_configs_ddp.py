"""BASELINE configs 4 and 5 on N GPUs of one node (torchrun), the way bench.py runs config 2: one process per GPU, the whole
step (forward, loss, backward with the bucketed NCCL gradient all-reduce overlapped, fused AdamW) replayed as one CUDA graph,
CUDA-event timing between barriers, max over ranks.  Prints one JSON line per config on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_configs_ddp.py --steps 20 --warmup 5 [c4] [c5]

config 4: ContrastiveEncoder(convnext_tiny, 2 ch, 15 slices) 2 views x 64 x (2,15,224,224) bf16 + NT-Xent(0.07), per-rank negatives
config 5: Unet3d(3,3,4,32) 1 x (3,128,128,128) fp16 autocast + GradScaler, MSE
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ContrastiveEncoder, Unet3d, _lib  # noqa: E402
from viscy_b200.graphs import GraphedStep  # noqa: E402
from viscy_b200.loss import NTXentLoss  # noqa: E402
from viscy_b200.parallel import BucketedGradAllReduce  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c4", "c5"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1
    if ddp:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    def run(name, model, make_step, inputs, samples_per_rank, tflop_per_rank, extra):
        exchange = None
        if ddp:
            exchange = BucketedGradAllReduce(model.parameters())
            exchange.broadcast_parameters(0)
        step = make_step(exchange)
        gs = GraphedStep(step, inputs, warmup=11 if ddp else 3)
        for _ in range(args.warmup):
            gs.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            gs.replay()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if ddp:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = ms.item() / args.steps
        loss = gs.static_output.float().item()
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": ms_step, "value": samples_per_rank * world / (ms_step * 1e-3),
                              "unit": "samples/s", "scaling": "weak", "algorithmic_tflops_per_gpu": tflop_per_rank / (ms_step * 1e-3),
                              "launches_per_step": gs.launches_per_replay, "loss": loss,
                              "grad_exchange": "bucketed NCCL all-reduce overlapped with backward" if ddp else "none", **extra}),
                  flush=True)
        del gs
        torch.cuda.empty_cache()

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if "c5" in args.which:
        torch.manual_seed(0)
        m = Unet3d(3, 3, 4, 32).to(dev)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True, capturable=True)
        scaler = torch.amp.GradScaler("cuda")
        x = torch.randn((1, 3, 128, 128, 128), device=dev, generator=g)
        y = torch.randn((1, 3, 128, 128, 128), device=dev, generator=g)

        def make5(exchange):
            def step5(xd, yd):
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = torch.nn.functional.mse_loss(m(xd).float(), yd)
                scaler.scale(loss).backward()
                if exchange is not None:
                    exchange.finish()
                scaler.step(opt)
                scaler.update()
                return loss
            return step5

        run("config5 Unet3d(3,3,4,32) 128^3 fp16 batch 1/GPU", m, make5, (x, y), 1, 3.696,
            {"dtype": "fp16", "amp": "fp16 autocast + GradScaler"})
        del m, opt, x, y
    if "c4" in args.which:
        torch.manual_seed(0)
        m = ContrastiveEncoder("convnext_tiny", in_channels=2, in_stack_depth=15).to(dev)
        opt = torch.optim.AdamW(m.parameters(), lr=2e-4, fused=True, capturable=True)
        a = torch.randn((64, 2, 15, 224, 224), device=dev, generator=g)
        p = torch.randn((64, 2, 15, 224, 224), device=dev, generator=g)
        labels = torch.cat([torch.arange(64), torch.arange(64)]).to(dev)
        crit = NTXentLoss(temperature=0.07)

        def make4(exchange):
            def step4(ad, pd):
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    _, pa = m(ad)
                    _, pp = m(pd)
                loss = crit(torch.cat([pa, pp]).float(), labels)
                loss.backward()
                if exchange is not None:
                    exchange.finish()
                opt.step()
                return loss
            return step4

        run("config4 ContrastiveEncoder convnext_tiny 2 views x 64/GPU 224x224x15 bf16 NT-Xent", m, make4, (a, p), 128, 3.448,
            {"dtype": "bf16"})
    if ddp:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
