import sys; sys.path.insert(0,"/root/repo")
import torch
from viscy_b200 import ops
dev=torch.device("cuda:0"); g=torch.Generator(device=dev).manual_seed(0)
for (B,R,C,w) in ((8,4096,2960,2944),(8,4096,400,384),(8,1024,784,768),(8,256,1552,1536),(8,64,3088,3072),(1,32768,736,736),(1,2097152,32,32)):
    x=torch.randn((B,R,C),device=dev,generator=g).bfloat16()
    f=lambda: ops.colreduce(x,1,width=w)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); us=e0.elapsed_time(e1)/20*1e3
    print(f"B={B} R={R} C={C}: {us:.1f} us  {x.numel()*2/us*1e-3:.0f} GB/s")
