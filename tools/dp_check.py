"""Data-parallel check on real GPUs (run under torchrun, world_size >= 2): the bucketed NCCL
all-reduce path (generator gradient sink + discriminator hooks) must give the same gradients as a
single-GPU run on the concatenated batch.  Dropout masks are injected so both runs see the same
masks.  Prints one line from rank 0 and exits non-zero on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import cunet_oracle as orc
from weather_unet_b200 import Conditional_UNet
from weather_unet_b200.disc import SNDisc
from weather_unet_b200.train_step import GDTrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per, H, nc = 2, 64, 5
Bg = per * world


def build():
    torch.manual_seed(0)
    G = Conditional_UNet(nc).to(dev).train()
    torch.manual_seed(100)
    D = SNDisc(nc).to(dev).train()
    return G, D


g = torch.Generator().manual_seed(3)
x = (torch.rand(Bg, 3, H, H, generator=g) * 2 - 1).to(dev)
cr = torch.eye(nc)[torch.randint(0, nc, (Bg,), generator=g)].to(dev)
ct = torch.eye(nc)[torch.randint(0, nc, (Bg,), generator=g)].to(dev)
md = orc.make_dropout_masks(Bg, H, H, seed=1, device=dev)
mg = orc.make_dropout_masks(Bg, H, H, seed=2, device=dev)
sl = slice(rank * per, (rank + 1) * per)

G, D = build()
step = GDTrainStep(G, D, lr=0.0)
assert step.distributed and step._use_sink
step.step(x[sl], cr[sl], ct[sl], masks_d=tuple(m[sl] for m in md), masks_g=tuple(m[sl] for m in mg))
torch.cuda.synchronize()
dp = {("G." + n): p.grad.detach().clone() for n, p in G.named_parameters() if p.grad is not None}
dp.update({("D." + n): p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None})
ok = True
if rank == 0:
    G1, D1 = build()
    single = GDTrainStep(G1, D1, lr=0.0, distributed=False)
    single.step(x, cr, ct, masks_d=md, masks_g=mg)
    torch.cuda.synchronize()
    ref = {("G." + n): p.grad for n, p in G1.named_parameters() if p.grad is not None}
    ref.update({("D." + n): p.grad for n, p in D1.named_parameters() if p.grad is not None})
    assert set(ref) == set(dp), set(ref) ^ set(dp)
    worst = ("", 0.0)
    for k in ref:
        r = ((dp[k].float() - ref[k].float()).norm() / (ref[k].float().norm() + 1e-20)).item()
        if r > worst[1]:
            worst = (k, r)
    # the all-reduced mean of per-shard gradients equals the full-batch gradient up to fp32
    # summation order (generator) / bf16 autocast noise (discriminator)
    ok = worst[1] < 2e-2
    print(f"dp_check world={world}: {len(ref)} gradients, worst rel-L2 {worst[1]:.3e} at {worst[0]} -> "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
