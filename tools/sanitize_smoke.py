"""Small invocations of every tcgen05 / mbarrier-pipelined kernel family and of the fused HBM-bound
kernels, sized so that `compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_smoke.py`
finishes in a minute or two (the 6-warp TMA / MMA / epilogue pipelines are where a race
would hide).  Prints one line per kernel family; the sanitizer's own summary is the result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import Conditional_UNet, _ops as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H = 1, 16


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


for cin, cout in ((64, 64), (64, 128), (128, 256), (192, 64)):
    c1 = 64 if cin == 192 else 0
    s0, s1 = bf(B, H, H, cin - c1), (bf(B, H, H, c1) if c1 else None)
    dy = bf(B, H, H, cout)
    wf, wd = K.pack_conv3x3_weights(torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    y = K.conv3x3(s0, s1, wf, bias, True, None, cout)
    if not c1:
        K.conv3x3(dy, None, wd, None, False, s0, cin)
    K.conv3x3_wgrad(s0, s1, dy)
    print(f"conv3x3 {cin}->{cout}: fprop / dgrad / wgrad launched", flush=True)
wf, _ = K.pack_conv3x3_weights(torch.randn(64, 64, 3, 3, device=dev) * 0.05)
K.conv3x3_pool(bf(B, H, H, 64), wf, torch.zeros(64, device=dev), 64)
K.conv3x3_last(bf(B, H, H, 64), wf, torch.zeros(64, device=dev), torch.randn(3, 64, 1, 1, device=dev),
               torch.zeros(3, device=dev))
print("fused pool / last epilogues launched", flush=True)
x = torch.rand(B, 3, H, H, device=dev) * 2 - 1
a1 = K.conv_first(x, torch.randn(64, 3, 3, 3, device=dev) * 0.2, torch.zeros(64, device=dev))
K.conv_first_wgrad(x, bf(B, H, H, 64))
print("K = 27 first-layer kernels launched", flush=True)
s2w, s2d = K.pack_conv3x3_weights(torch.randn(128, 64, 3, 3, device=dev) * 0.05)
y2 = K.conv3x3_s2(bf(B, H, H, 64), s2w, torch.zeros(128, device=dev), 0.2, 128)
K.conv3x3_s2_dgrad(bf(B, H // 2, H // 2, 128), s2d, 64, H, H)
K.conv3x3_s2_wgrad(bf(B, H, H, 64), bf(B, H // 2, H // 2, 128))
print("stride-2 kernels launched", flush=True)
G = Conditional_UNet(5).to(dev).train()
c = torch.eye(5, device=dev)[:1]
y = G(torch.rand(1, 3, 16, 16, device=dev) * 2 - 1, c, seed=3)
y.mean().backward()
torch.cuda.synchronize()
print("generator forward + backward (AdaIN / upsample / dropout / pooling kernels) done", flush=True)
