"""One G+D training iteration bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off.
Usage: python tools/profile_step.py [batch] [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import Conditional_UNet
from weather_unet_b200.disc import SNDisc
from weather_unet_b200.train_step import GDTrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
G = Conditional_UNet(5).to(dev).train()
torch.manual_seed(100)
D = SNDisc(5).to(dev).train()
tr = GDTrainStep(G, D)
g = torch.Generator().manual_seed(1)
img = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev)
cr = torch.eye(5)[torch.randint(0, 5, (B,), generator=g)].to(dev)
ct = torch.eye(5)[torch.randint(0, 5, (B,), generator=g)].to(dev)
for _ in range(3):
    tr.step(img, cr, ct)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = tr.step(img, cr, ct)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: float(v) for k, v in out.items()})
