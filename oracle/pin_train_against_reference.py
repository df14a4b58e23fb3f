"""Pins oracle/train_oracle.py (discriminator + one G+D iteration) against the reference modules
imported from /root/reference, and writes tests/golden/train_b2_h32_seed0.npz: the loss curve of
N iterations at fixed seed with the dropout masks drawn from the seeded torch RNG stream.
Build-container only.  Run:  python -O oracle/pin_train_against_reference.py
"""
import os
import subprocess
import sys

if __debug__:
    sys.exit(subprocess.call([sys.executable, "-O"] + sys.argv))

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import train_oracle as T  # noqa: E402
from oracle.pin_against_reference import load_ref  # noqa: E402

N_STEPS, LR = 6, 1e-3


def ref_iteration(Gm, Dm, g_opt, d_opt, images, c_real, c_target):
    """The reference's update_discriminator + update_inference (supervised branch, estimator term
    omitted), driven with the reference's own nn.Modules."""
    d_opt.zero_grad()
    real = Dm(images, c_real)[0]
    fake_img = Gm(images, c_target)
    fake = Dm(fake_img.detach(), c_target)[0]
    d_loss = torch.mean(torch.relu(1. - real)) + torch.mean(torch.relu(1. + fake))
    d_loss.backward()
    d_opt.step()
    g_opt.zero_grad()
    fake_img = Gm(images, c_target)
    fake = Dm(fake_img, c_target)[0]
    g_adv = torch.mean(-fake)
    g_l1 = F.l1_loss(fake_img, images)
    diff = torch.mean(torch.abs(fake_img - images), [1, 2, 3])
    lmda = torch.mean(torch.abs(c_real - c_target), 1)
    loss_con = torch.mean(diff / (lmda + 1e-2))
    g_loss = g_adv + loss_con
    g_loss.backward()
    g_opt.step()
    return {"d_loss": d_loss.item(), "g_loss": g_loss.item(), "g_loss_adv": g_adv.item(),
            "g_loss_l1": g_l1.item(), "loss_con": loss_con.item()}


def main():
    torch.set_num_threads(4)
    ref_cunet, ref_disc = load_ref("cunet"), load_ref("disc")
    nc, B, H = 5, 2, 32
    torch.manual_seed(0)
    Gm = ref_cunet.Conditional_UNet(nc)
    torch.manual_seed(100)
    Dm = ref_disc.SNDisc(nc)
    g_sd = {k: v.clone() for k, v in Gm.state_dict().items()}
    d_sd = {k: v.clone() for k, v in Dm.state_dict().items()}

    # our SNDisc restatement draws the same init from the same seed (checked here, relied on by tests)
    import weather_unet_b200.disc as our_disc
    torch.manual_seed(100)
    ours = our_disc.SNDisc(nc)
    for k, v in ours.state_dict().items():
        assert torch.equal(v, d_sd[k]), f"SNDisc init differs at {k}"
    assert list(ours.state_dict().keys()) == list(d_sd.keys())

    g = torch.Generator().manual_seed(2)
    images = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    c_real = torch.eye(nc)[torch.randint(0, nc, (B,), generator=g)]
    c_target = torch.eye(nc)[torch.randint(0, nc, (B,), generator=g)]

    # discriminator forward alone: bit-exact, including the power-iteration buffer updates
    d_tmp = {k: v.clone() for k, v in d_sd.items()}
    Dm.train()
    out_ref = Dm(images, c_real)
    out_orc = T.disc_forward(d_tmp, images, c_real, train=True)
    for a, b in zip(out_ref, out_orc):
        assert torch.equal(a, b), "discriminator forward differs"
    for k, v in Dm.state_dict().items():
        if k.endswith("_u") or k.endswith("_v"):
            assert torch.equal(v, d_tmp[k]), f"power iteration buffer differs at {k}"
    print("SNDisc forward + u/v update: oracle == reference bit for bit")
    Dm.load_state_dict(d_sd)

    g_opt = torch.optim.Adam(Gm.parameters(), lr=LR, betas=(0.0, 0.999), weight_decay=LR / 20)
    d_opt = torch.optim.Adam(Dm.parameters(), lr=LR, betas=(0.0, 0.999), weight_decay=LR / 20)
    trainer = T.Trainer(g_sd, d_sd, lr=LR)
    Gm.train(), Dm.train()
    curve_ref, curve_orc = [], []
    torch.manual_seed(11)
    for _ in range(N_STEPS):
        curve_ref.append(ref_iteration(Gm, Dm, g_opt, d_opt, images, c_real, c_target))
    torch.manual_seed(11)
    for _ in range(N_STEPS):
        curve_orc.append(trainer.step(images, c_real, c_target))
    keys = sorted(curve_ref[0])
    ref_arr = np.array([[s[k] for k in keys] for s in curve_ref])
    orc_arr = np.array([[s[k] for k in keys] for s in curve_orc])
    err = np.abs(ref_arr - orc_arr).max()
    print("loss curve (reference modules):")
    print(keys)
    print(ref_arr)
    assert err < 1e-5, f"loss curves differ by {err}"
    print(f"{N_STEPS}-iteration loss curve: oracle vs reference modules max |diff| = {err:.2e}")
    for k, v in Gm.state_dict().items():
        assert torch.allclose(v, trainer.g[k].detach(), atol=1e-5), k

    # masks used by each generator forward, re-drawn from the same stream, for the GPU loss-curve test
    torch.manual_seed(11)
    masks = []
    for _ in range(2 * N_STEPS):
        ms = []
        for shape in ((B, 512, H // 4, H // 4), (B, 256, H // 2, H // 2), (B, 128, H, H)):
            keep = torch.empty(shape).bernoulli_(0.7)
            ms.append(np.packbits(keep.permute(0, 2, 3, 1).to(torch.uint8).numpy().reshape(-1)))
        masks.append(ms)
    out = {"images": images.numpy(), "c_real": c_real.numpy(), "c_target": c_target.numpy(),
           "keys": np.array(keys), "curve": ref_arr, "lr": np.array([LR]),
           "d_checksum": np.array([[v.double().sum().item(), v.double().abs().sum().item()]
                                   for v in d_sd.values()])}
    for i, ms in enumerate(masks):
        for j, m in enumerate(ms):
            out[f"mask_{i}_{j}"] = m
    path = os.path.join(ROOT, "tests", "golden", "train_b2_h32_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
