"""CUDA-event timings of the AdaIN + upsample + dropout kernels at the three decoder sites of the
bench step (batch 64, 256x256 images): forward kernel alone (Philox dropout), forward incl. statistics
and style, and the backward (adjoint + style + apply).  Usage: python tools/time_adain.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weather_unet_b200 import _lib
if os.environ.get("WU_TIME_LIB"):  # A/B against another build of the library
    _lib.LIB_PATH = os.environ["WU_TIME_LIB"]
from weather_unet_b200 import _ops as K
from weather_unet_b200._ops import call, ptr, stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
cond = torch.eye(5, device=dev)[torch.randint(0, 5, (B,))]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = [0.0, 0.0, 0.0]
for h, C in ((32, 512), (64, 256), (128, 128)):
    x = torch.randn(B, h, h, C, device=dev).clamp_min(0).to(torch.bfloat16)
    lw = torch.randn(4 * C, 5, device=dev) * 0.3
    lb = torch.zeros(4 * C, device=dev)
    gu = torch.randn(B, 2 * h, 2 * h, C, device=dev).to(torch.bfloat16)
    u, st = K.adain_up_drop(x, cond, lw, lb, 1e-5, 0.3, 1234, None)

    def fwd_only():
        call("wu_adain_up_drop_fwd", ptr(x), ptr(st.scale), ptr(st.shift), ptr(u), ptr(st.bits), B, h, h,
             C, 0.3, 1234, ptr(None), 0, stream())

    t0 = timed(fwd_only)
    t1 = timed(lambda: K.adain_up_drop(x, cond, lw, lb, 1e-5, 0.3, 1234, None))
    t2 = timed(lambda: K.adain_up_drop_bwd(gu, x, cond, lw, lb, st))
    out_bytes = u.numel() * 2 + st.bits.numel() + x.numel() * 2
    print(f"site {h}x{h}x{C}: fwd kernel {t0:.4f} ms ({out_bytes / t0 / 1e6:.0f} GB/s algorithmic), "
          f"fwd incl. stats+style {t1:.4f} ms, backward {t2:.4f} ms")
    tot = [tot[0] + t0, tot[1] + t1, tot[2] + t2]
print(f"all three sites: fwd kernel {tot[0]:.4f} ms, fwd total {tot[1]:.4f} ms, backward {tot[2]:.4f} ms")
