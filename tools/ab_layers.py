"""A/B timing of two builds of the library IN ONE PROCESS (box-to-box and minute-to-minute clock
differences are larger than most kernel changes): for every generator layer, fprop / dgrad / wgrad
are timed alternately on build A and build B (CUDA events, several rounds, medians).
Usage: python tools/ab_layers.py <libA.so> <libB.so> [B] [H]"""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from weather_unet_b200 import _lib
from weather_unet_b200 import _ops as K


def load_lib(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


libs = {"A": load_lib(sys.argv[1]), "B": load_lib(sys.argv[2])}
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
H = int(sys.argv[4]) if len(sys.argv) > 4 else 256
dev = torch.device("cuda:0")
layers = [("down1.2", 64, 0, 64, 1), ("down2.0", 64, 0, 128, 2), ("down2.2", 128, 0, 128, 2),
          ("down3.0", 128, 0, 256, 4), ("down3.2", 256, 0, 256, 4), ("down4.0", 256, 0, 512, 8),
          ("down4.2", 512, 0, 512, 8), ("up3.0", 512, 256, 256, 4), ("up3.2", 256, 0, 256, 4),
          ("up2.0", 256, 128, 128, 2), ("up2.2", 128, 0, 128, 2), ("up1.0", 128, 64, 64, 1),
          ("up1.2", 64, 0, 64, 1)]
if len(sys.argv) > 5:
    layers = [l for l in layers if l[0] in sys.argv[5:]]


def time_ms(fn, it=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


tot = {k: {"A": 0.0, "B": 0.0} for k in ("fprop", "dgrad", "wgrad")}
for name, c0, c1, cout, d in layers:
    h = H // d
    s0 = torch.randn(B, h, h, c0, device=dev).to(torch.bfloat16)
    s1 = torch.randn(B, h, h, c1, device=dev).to(torch.bfloat16) if c1 else None
    dy = torch.randn(B, h, h, cout, device=dev).to(torch.bfloat16)
    _lib._lib = libs["A"]
    wf, wd = K.pack_conv3x3_weights(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    fns = {"fprop": lambda: K.conv3x3(s0, s1, wf, bias, True, None, cout),
           "dgrad": (lambda: (K.conv3x3(dy, None, wd[:c0], None, False, None, c0),
                              K.conv3x3(dy, None, wd[c0:], None, False, None, c1))) if c1 else
                    (lambda: K.conv3x3(dy, None, wd, None, False, s0, c0)),
           "wgrad": lambda: K.conv3x3_wgrad(s0, s1, dy)}
    line = f"{name:8s}"
    for kind, fn in fns.items():
        t = {"A": [], "B": []}
        for rnd in range(5):
            for which in ("A", "B") if rnd % 2 == 0 else ("B", "A"):
                _lib._lib = libs[which]
                t[which].append(time_ms(fn))
        ma, mb = statistics.median(t["A"]), statistics.median(t["B"])
        tot[kind]["A"] += ma
        tot[kind]["B"] += mb
        line += f" | {kind} A {ma:.4f} B {mb:.4f} ms ({100 * (ma / mb - 1):+5.1f}% speed-up of B)"
    print(line, flush=True)
for kind, v in tot.items():
    print(f"TOTAL {kind}: A {v['A']:.3f} ms, B {v['B']:.3f} ms ({100 * (v['A'] / v['B'] - 1):+.1f}% speed-up of B)")
