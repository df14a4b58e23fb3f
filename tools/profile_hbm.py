"""Launch each HBM / instruction-bound kernel of the step once at bench shapes (B=64, 256x256) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import _ops as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.manual_seed(0)
img = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
w1 = torch.randn(64, 3, 3, 3, device=dev) * 0.2
b1 = torch.zeros(64, device=dev)
dy256 = torch.randn(B, 256, 256, 64, device=dev).to(bf)
x128 = torch.randn(B, 128, 128, 128, device=dev).to(bf)
cond = torch.eye(5, device=dev)[torch.randint(0, 5, (B,))]
lw = torch.randn(512, 5, device=dev) * 0.3
lb = torch.zeros(512, device=dev)
gu = torch.randn(B, 256, 256, 128, device=dev).to(bf)
y256 = torch.randn(B, 256, 256, 64, device=dev).to(bf)
gp = torch.randn(B, 128, 128, 64, device=dev).to(bf)


# discriminator pieces (disc.py:28-31 at batch 64): stem + first trunk block, forward and backward
w0s = (torch.randn(3, 3, 3, 3, device=dev) * 0.3).requires_grad_(True)
b0s = torch.zeros(3, device=dev, requires_grad=True)
w1s = (torch.randn(64, 3, 3, 3, device=dev) * 0.2).requires_grad_(True)
b1s = torch.zeros(64, device=dev, requires_grad=True)
wa = (torch.randn(64, 64, 3, 3, device=dev) * 0.04).requires_grad_(True)
ba = torch.zeros(64, device=dev, requires_grad=True)
wb = (torch.randn(128, 64, 3, 3, device=dev) * 0.04).requires_grad_(True)
bb = torch.zeros(128, device=dev, requires_grad=True)
img_g = img.clone().requires_grad_(True)


def disc_part():
    c1 = K.disc_stem(img_g, w0s, b0s, w1s, b1s, 0.2)
    c2 = K.disc_block(c1, wa, ba, wb, bb, 0.2)
    c2.float().sum().backward()


def run():
    disc_part()
    K.conv_first(img, w1, b1)
    K.conv_first_wgrad(img, dy256)
    u, st = K.adain_up_drop(x128, cond, lw, lb, 1e-5, 0.3, 1234, None)
    K.adain_up_drop_bwd(gu, x128, cond, lw, lb, st)
    K.maxpool2(y256)
    K.maxpool2_bwd(y256, gp, dy256)
    K.conv3x3_wgrad(y256, None, dy256)  # includes bias_grad_partial / wgrad_reduce


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
