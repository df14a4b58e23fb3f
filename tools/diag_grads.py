"""Per-activation / per-parameter error report of the CUDA generator vs the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cunet_oracle as orc
from weather_unet_b200 import Conditional_UNet

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, H, W, nc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[2]), 5
train = len(sys.argv) > 3 and sys.argv[3] == "train"
torch.manual_seed(3)
net = Conditional_UNet(nc).to(dev).train(train)
g = torch.Generator().manual_seed(B + H)
x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
c = torch.randn(B, nc, generator=g).to(dev)
gy = torch.randn(B, 3, H, W, generator=g).to(dev)
masks = orc.make_dropout_masks(B, H, W, seed=5, device=dev) if train else None
acts = {}
y = net(x, c, dropout_masks=masks, _keep_acts=acts)
(y * gy).sum().backward()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
col = {}
leaf = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
y_ref = orc.forward(leaf, x, c, train=train, masks=masks, collect=col)
(y_ref * gy).sum().backward()
# autocast-bf16 comparator: the error band a stock bf16 PyTorch run has against fp32
leaf2 = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    y_ac = orc.forward(leaf2, x, c, train=train, masks=masks)
(y_ac.float() * gy).sum().backward()
print(f"B={B} H={H} train={train}")
for k in ("conv1", "conv2", "conv3", "x4", "u3", "up3b", "u2", "up2b", "u1", "up1b"):
    a = acts[k].float().permute(0, 3, 1, 2)
    r = col[k]
    print(f"act {k:6s} rel-L2 {((a - r).norm() / r.norm()).item():.3e}")
print(f"y max-abs {(y.detach() - y_ref.detach()).abs().max().item():.3e}   autocast-bf16: {(y_ac.float().detach() - y_ref.detach()).abs().max().item():.3e}")
for name, p in net.named_parameters():
    if p.grad is None:
        continue
    gm, gr, ga = p.grad.float().flatten(), leaf[name].grad.flatten(), leaf2[name].grad.float().flatten()
    r = ((gm - gr).norm() / gr.norm()).item()
    ra = ((ga - gr).norm() / gr.norm()).item()
    cos = torch.nn.functional.cosine_similarity(gm, gr, dim=0).item()
    print(f"grad {name:24s} rel-L2 {r:.3e} cos {cos:.5f} | autocast-bf16 rel-L2 {ra:.3e}")
