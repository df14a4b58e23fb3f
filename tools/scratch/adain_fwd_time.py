import sys, torch
sys.path.insert(0, ".")
from weather_unet_b200 import _ops as K
dev = torch.device("cuda:0")
B = 64
for (C, h) in ((128, 128), (256, 64), (512, 32)):
    x = torch.randn(B, h, h, C, device=dev).to(torch.bfloat16)
    cond = torch.eye(5, device=dev)[torch.randint(0, 5, (B,))]
    lw, lb = torch.randn(4 * C, 5, device=dev) * 0.3, torch.zeros(4 * C, device=dev)
    f = lambda: K.adain_up_drop(x, cond, lw, lb, 1e-5, 0.3, 7, None)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    print(C, h, round(a.elapsed_time(b) / 10, 4), "ms")
