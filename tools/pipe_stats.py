"""Where do the single MMA-issuing thread and the TMA producer of the 3x3 convolution kernels wait?
Runs each generator layer (batch 64, 256x256 by default) on the -DWU_PIPE_STATS debug build
(tools/pipe_stats.sh) and prints the share of each thread's loop spent in each barrier wait."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from weather_unet_b200 import _lib

_lib.LIB_PATH = os.environ.get("WU_B200_LIB") or os.path.join(ROOT, "tools", "scratch", "libwu_b200_stats.so")
from weather_unet_b200 import _ops as K  # noqa: E402

lib = _lib.load()
lib.wu_debug_pipe_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
layers = [("down1.2", 64, 0, 64, 1), ("down2.0", 64, 0, 128, 2), ("down2.2", 128, 0, 128, 2),
          ("down3.0", 128, 0, 256, 4), ("down3.2", 256, 0, 256, 4), ("down4.2", 512, 0, 512, 8),
          ("up3.0", 512, 256, 256, 4), ("up2.0", 256, 128, 128, 2), ("up2.2", 128, 0, 128, 2),
          ("up1.0", 128, 64, 64, 1)]


def stats(fn):
    fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib.wu_debug_pipe_stats(None, 1)
    fn()
    torch.cuda.synchronize()
    lib.wu_debug_pipe_stats(buf, 1)
    return list(buf)


def pct(a, b):
    return f"{100.0 * a / b:5.1f}%" if b else "   - "


for name, c0, c1, cout, d in layers:
    h = S // d
    s0 = torch.randn(B, h, h, c0, device=dev).to(torch.bfloat16)
    s1 = torch.randn(B, h, h, c1, device=dev).to(torch.bfloat16) if c1 else None
    dy = torch.randn(B, h, h, cout, device=dev).to(torch.bfloat16)
    wf, wd = K.pack_conv3x3_weights(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    f = stats(lambda: K.conv3x3(s0, s1, wf, bias, True, None, cout))
    w = stats(lambda: K.conv3x3_wgrad(s0, s1, dy))
    line = f"{name:8s} {c0 + c1:4d}->{cout:3d} @{h:3d}^2 | fprop"
    if cout % 256 == 0:  # v1 kernel: one barrier per (A, B) stage
        line += (f" v1 MMA thread waits: tmem-empty {pct(f[1], f[0])} full {pct(f[2], f[0])}"
                 f" | producer waits: empty {pct(f[5], f[4])}")
    elif f[0]:
        line += (f" MMA thread waits: tmem-empty {pct(f[1], f[0])} fullA {pct(f[2], f[0])} fullB {pct(f[3], f[0])}"
                 f" | producer waits: emptyA {pct(f[5], f[4])} emptyB {pct(f[6], f[4])}")
    else:
        line += " (?)"
    line += f" | wgrad MMA waits full {pct(w[9], w[8])}, producer waits empty {pct(w[13], w[12])}"
    print(line, flush=True)
    if not c1 and f[0]:
        g = stats(lambda: K.conv3x3(dy, None, wd, None, False, s0, c0))
        if g[0]:
            print(f"{'':27s} | dgrad MMA thread waits: tmem-empty {pct(g[1], g[0])} fullA {pct(g[2], g[0])} "
                  f"fullB {pct(g[3], g[0])} | producer waits: emptyA {pct(g[5], g[4])} emptyB {pct(g[6], g[4])}")
