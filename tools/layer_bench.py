"""Per-layer timing of the tensor-core convolution kernels (fprop / dgrad / wgrad) at the
generator's shapes.  Usage: python tools/layer_bench.py [B] [H] [iters]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import _ops as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
PEAK = 1661.6
layers = [  # name, c0, c1, cout, downscale
    ("down1.2", 64, 0, 64, 1), ("down2.0", 64, 0, 128, 2), ("down2.2", 128, 0, 128, 2),
    ("down3.0", 128, 0, 256, 4), ("down3.2", 256, 0, 256, 4), ("down4.0", 256, 0, 512, 8),
    ("down4.2", 512, 0, 512, 8), ("up3.0", 512, 256, 256, 4), ("up3.2", 256, 0, 256, 4),
    ("up2.0", 256, 128, 128, 2), ("up2.2", 128, 0, 128, 2), ("up1.0", 128, 64, 64, 1),
    ("up1.2", 64, 0, 64, 1)]


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


tot = {"fprop": [0, 0], "dgrad": [0, 0], "wgrad": [0, 0]}
rows = []
for name, c0, c1, cout, d in layers:
    h = H // d
    cin = c0 + c1
    s0 = torch.randn(B, h, h, c0, device=dev).to(torch.bfloat16)
    s1 = torch.randn(B, h, h, c1, device=dev).to(torch.bfloat16) if c1 else None
    dy = torch.randn(B, h, h, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    bias = torch.zeros(cout, device=dev)
    wf, wd = K.pack_conv3x3_weights(w)
    fl = 2 * 9 * cin * cout * h * h * B
    t_f = timeit(lambda: K.conv3x3(s0, s1, wf, bias, True, None, cout))
    if c1:
        t_d = timeit(lambda: (K.conv3x3(dy, None, wd[:c0], None, False, None, c0),
                              K.conv3x3(dy, None, wd[c0:], None, False, None, c1)))
    else:
        t_d = timeit(lambda: K.conv3x3(dy, None, wd, None, False, s0, cin))
    t_w = timeit(lambda: K.conv3x3_wgrad(s0, s1, dy))
    for k, t in (("fprop", t_f), ("dgrad", t_d), ("wgrad", t_w)):
        tot[k][0] += fl
        tot[k][1] += t
    rows.append((name, cin, cout, h, fl / 1e9, t_f, t_d, t_w))
    print(f"{name:8s} cin {cin:4d} cout {cout:4d} {h:4d}^2  {fl/1e9:8.1f} GF | fprop {t_f:7.3f} ms {fl/t_f/1e9:7.1f} TF/s"
          f" ({fl/t_f/1e9/PEAK*100:4.1f}%) | dgrad {t_d:7.3f} ms {fl/t_d/1e9:7.1f} TF/s | wgrad {t_w:7.3f} ms {fl/t_w/1e9:7.1f} TF/s",
          flush=True)
for k, (fl, t) in tot.items():
    print(f"TOTAL {k}: {fl/1e9:.1f} GF in {t:.3f} ms = {fl/t/1e9:.1f} TF/s ({fl/t/1e9/PEAK*100:.1f}% of {PEAK} measured bf16 peak)")
