"""Pins a 50-iteration fixed-seed training run (N = 50) of the oracle trainer
(oracle/train_oracle.py) against the reference's own nn.Modules imported from /root/reference and
writes tests/golden/train50_b2_h32_seed0.npz: the loss curve and, per parameter tensor of G and D,
how far the 50 updates moved it.  At lr 1e-3 the curve moves by more than half (loss_con 26.2 ->
10.5), so a generator that does not actually train cannot reproduce it.

The dropout masks are NOT stored: every generator forward draws three masks from the seeded CPU
RNG stream exactly like nn.Dropout does (torch.empty(shape).bernoulli_(0.7), NCHW shapes, sites
adain3 / adain2 / adain1 in that order); `draw_masks` below is that recipe and the GPU test
imports it.  Build-container only.  Run:  python -O oracle/pin_train50_against_reference.py
"""
import os
import subprocess
import sys

if __debug__ and __name__ == "__main__":
    sys.exit(subprocess.call([sys.executable, "-O"] + sys.argv))

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_STEPS, LR, SEED_MASKS = 50, 1e-3, 11


def draw_masks(B, H, W, gen=None):
    """Three keep-masks of one generator forward in the order / shapes nn.Dropout draws them
    (cunet.py:61,68,75), returned as uint8 NHWC (what both Conditional_UNet(dropout_masks=...) and
    the oracle's forward(masks=...) take)."""
    out = []
    for shape in ((B, 512, H // 4, W // 4), (B, 256, H // 2, W // 2), (B, 128, H, W)):
        keep = torch.empty(shape).bernoulli_(0.7, generator=gen)
        out.append(keep.permute(0, 2, 3, 1).to(torch.uint8).contiguous())
    return tuple(out)


def tensor_stats(final, init):
    """[|final|, |final - init|, sum(final)] in float64."""
    f, i = final.detach().double().cpu(), init.detach().double().cpu()
    return [f.norm().item(), (f - i).norm().item(), f.sum().item()]


def main():
    from oracle import train_oracle as T
    from oracle.pin_against_reference import load_ref
    from oracle.pin_train_against_reference import ref_iteration
    torch.set_num_threads(4)
    ref_cunet, ref_disc = load_ref("cunet"), load_ref("disc")
    nc, B, H = 5, 2, 32
    torch.manual_seed(0)
    Gm = ref_cunet.Conditional_UNet(nc)
    torch.manual_seed(100)
    Dm = ref_disc.SNDisc(nc)
    g_sd = {k: v.clone() for k, v in Gm.state_dict().items()}
    d_sd = {k: v.clone() for k, v in Dm.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    images = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    c_real = torch.eye(nc)[torch.randint(0, nc, (B,), generator=g)]
    c_target = torch.eye(nc)[torch.randint(0, nc, (B,), generator=g)]

    g_opt = torch.optim.Adam(Gm.parameters(), lr=LR, betas=(0.0, 0.999), weight_decay=LR / 20)
    d_opt = torch.optim.Adam(Dm.parameters(), lr=LR, betas=(0.0, 0.999), weight_decay=LR / 20)
    Gm.train(), Dm.train()
    torch.manual_seed(SEED_MASKS)  # the reference draws its dropout masks from the global stream
    curve_ref = [ref_iteration(Gm, Dm, g_opt, d_opt, images, c_real, c_target) for _ in range(N_STEPS)]

    # the oracle with the masks injected from the same stream (the recipe the GPU test uses)
    trainer = T.Trainer(g_sd, d_sd, lr=LR)
    torch.manual_seed(SEED_MASKS)
    curve_orc = []
    for _ in range(N_STEPS):
        md, mg = draw_masks(B, H, H), draw_masks(B, H, H)
        curve_orc.append(trainer.step(images, c_real, c_target, masks_d=md, masks_g=mg))
    keys = sorted(curve_ref[0])
    ref_arr = np.array([[s[k] for k in keys] for s in curve_ref])
    orc_arr = np.array([[s[k] for k in keys] for s in curve_orc])
    err = np.abs(ref_arr - orc_arr).max()
    print(keys)
    print(ref_arr[::7])
    assert err < 1e-4, f"loss curves differ by {err}"
    print(f"{N_STEPS}-iteration loss curve: oracle (injected masks) vs reference modules max |diff| = {err:.2e}")
    worst = 0.0
    for k, v in Gm.state_dict().items():
        worst = max(worst, (v - trainer.g[k].detach()).abs().max().item())
    for k, v in Dm.state_dict().items():
        worst = max(worst, (v - trainer.d[k].detach()).abs().max().item())
    assert worst < 1e-4, worst
    print(f"final G and D state: oracle vs reference modules max |diff| = {worst:.2e}")

    out = {"images": images.numpy(), "c_real": c_real.numpy(), "c_target": c_target.numpy(),
           "keys": np.array(keys), "curve": ref_arr, "lr": np.array([LR]),
           "seed_masks": np.array([SEED_MASKS]),
           "g_names": np.array(list(g_sd.keys())), "d_names": np.array(list(d_sd.keys())),
           "g_stats": np.array([tensor_stats(Gm.state_dict()[k], g_sd[k]) for k in g_sd]),
           "d_stats": np.array([tensor_stats(Dm.state_dict()[k], d_sd[k]) for k in d_sd])}
    path = os.path.join(ROOT, "tests", "golden", "train50_b2_h32_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
