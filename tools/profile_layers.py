"""Launch a handful of representative convolution layers once each (after a warm-up) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`.  B=64, 256x256 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import _ops as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
cases = [("dconv_up1.0", 128, 64, 64, 256), ("dconv_down2.2", 128, 0, 128, 128), ("dconv_up3.0", 512, 256, 256, 64)]
if len(sys.argv) > 2:  # restrict to the named layers
    cases = [c for c in cases if c[0] in sys.argv[2:]]
work = []
for name, c0, c1, cout, h in cases:
    s0 = torch.randn(B, h, h, c0, device=dev).to(torch.bfloat16)
    s1 = torch.randn(B, h, h, c1, device=dev).to(torch.bfloat16) if c1 else None
    dy = torch.randn(B, h, h, cout, device=dev).to(torch.bfloat16)
    wf, wd = K.pack_conv3x3_weights(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    work.append((name, s0, s1, dy, wf, wd, bias, cout))


def run():
    for name, s0, s1, dy, wf, wd, bias, cout in work:
        K.conv3x3(s0, s1, wf, bias, True, None, cout)          # fprop
        K.conv3x3_wgrad(s0, s1, dy)                            # wgrad (+ reduce, bias grad)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
