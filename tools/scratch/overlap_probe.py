"""Does an instruction-bound elementwise kernel overlap with a persistent tensor-core conv kernel when
they are launched on two streams?  Prints serial vs concurrent wall time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from weather_unet_b200 import _ops as K
dev = torch.device("cuda:0")
B = 64
bf = torch.bfloat16
x128 = torch.randn(B, 128, 128, 128, device=dev).to(bf)
cond = torch.eye(5, device=dev)[torch.randint(0, 5, (B,))]
lw, lb = torch.randn(512, 5, device=dev) * 0.3, torch.zeros(512, device=dev)
cases = {"N=64 (192->64 @256^2)": (128, 64, 64, 256), "N=256 (768->256 @64^2)": (512, 256, 256, 64)}
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, (c0, c1, cout, h) in cases.items():
    a0 = torch.randn(B, h, h, c0, device=dev).to(bf)
    a1 = torch.randn(B, h, h, c1, device=dev).to(bf)
    wf, _ = K.pack_conv3x3_weights(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05)
    bias = torch.zeros(cout, device=dev)
    conv = lambda: K.conv3x3(a0, a1, wf, bias, True, None, cout)
    elem = lambda: K.adain_up_drop(x128, cond, lw, lb, 1e-5, 0.3, 7, None)

    def timed(fn, it=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it

    def both():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            conv()
        with torch.cuda.stream(s2):
            elem()
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    tc, te, tb = timed(conv), timed(elem), timed(both)
    print(f"{name}: conv {tc:.3f} ms, elementwise {te:.3f} ms, serial {tc + te:.3f} ms, two streams {tb:.3f} ms")
