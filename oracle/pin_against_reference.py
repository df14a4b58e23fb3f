"""Pins oracle/cunet_oracle.py against the UNMODIFIED reference imported from /root/reference and
writes the golden fixtures under tests/golden/.  Runs in the build container only (the GPU box has
no /root/reference).  Run:  python -O oracle/pin_against_reference.py
(-O strips the reference's `assert isinstance(x, torch.cuda.FloatTensor)`, utils.py:35, which
otherwise forbids CPU tensors; nothing else about the reference is changed.)
"""
import os
import subprocess
import sys

if __debug__:  # re-exec under -O
    sys.exit(subprocess.call([sys.executable, "-O"] + sys.argv))

import importlib.util

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import cunet_oracle as orc  # noqa: E402


def load_ref(name):
    """Import a reference module by path under a private name (ours are called the same)."""
    sys.path.insert(0, REF)
    try:
        for m in ("utils", "nets", "cunet", "disc", "ops"):
            sys.modules.pop(m, None)
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.remove(REF)
        for m in ("utils", "nets"):
            sys.modules.pop(m, None)


def main():
    torch.set_num_threads(4)
    torch.use_deterministic_algorithms(True)
    ref_cunet = load_ref("cunet")
    nc, B, H = 5, 2, 32
    torch.manual_seed(0)
    ref = ref_cunet.Conditional_UNet(nc)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}

    # our module must draw the same initial weights from the same seed (same construction order)
    import weather_unet_b200
    torch.manual_seed(0)
    ours = weather_unet_b200.Conditional_UNet(nc)
    osd = ours.state_dict()
    assert list(osd.keys()) == list(sd.keys()), "state_dict key order differs"
    for k in sd:
        assert torch.equal(sd[k], osd[k]), f"init differs at {k}"
    ours.load_state_dict(sd, strict=True)
    ref.load_state_dict(osd, strict=True)
    n_params = sum(v.numel() for v in sd.values())
    print(f"state_dict: {len(sd)} tensors, {n_params} params — keys, order and seeded init identical")

    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    c_hot = torch.eye(nc)[torch.randint(0, nc, (B,), generator=g)]
    c_soft = torch.randn(B, nc, generator=g)
    gy = torch.randn(B, 3, H, H, generator=g)
    out = {"x": x.numpy(), "c_hot": c_hot.numpy(), "c_soft": c_soft.numpy(), "gy": gy.numpy()}

    # eval mode: bit-exact
    ref.eval()
    with torch.no_grad():
        for tag, c in (("hot", c_hot), ("soft", c_soft)):
            y_ref = ref(x, c)
            y_orc = orc.forward(sd, x, c, train=False)
            assert torch.equal(y_ref, y_orc), f"eval forward differs ({tag})"
            out[f"y_eval_{tag}"] = y_ref.numpy()
    print("eval forward: oracle == reference bit for bit")

    # train mode: same RNG stream -> same dropout masks -> bit-exact forward and gradients
    ref.train()
    for tag, c in (("hot", c_hot), ("soft", c_soft)):
        ref.zero_grad()
        torch.manual_seed(7)
        y_ref = ref(x, c)
        (y_ref * gy).sum().backward()
        torch.manual_seed(7)
        col = {}
        leaf = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        y_orc = orc.forward(leaf, x, c, train=True, collect=col)
        (y_orc * gy).sum().backward()
        assert torch.equal(y_ref, y_orc), f"train forward differs ({tag})"
        live = 0
        for name, p in ref.named_parameters():
            if p.grad is None:
                assert leaf[name].grad is None and name.endswith("emb.weight"), name
                continue
            live += 1
            assert torch.equal(p.grad, leaf[name].grad), f"grad differs at {name} ({tag})"
        assert live == 36
        # the same masks injected explicitly give the same result again
        masks = tuple(col[f"mask{i}"].permute(0, 2, 3, 1).to(torch.uint8).contiguous() for i in (3, 2, 1))
        y_inj, g_inj = orc.forward_backward(sd, x, c, masks, gy)
        assert torch.equal(y_inj, y_ref)
        out[f"y_train_{tag}"] = y_ref.detach().numpy()
        if tag == "hot":
            for i, m in zip((3, 2, 1), masks):
                out[f"mask{i}_bits"] = np.packbits(m.numpy().reshape(-1))
                out[f"mask{i}_shape"] = np.array(m.shape)
            for k in ("conv1", "conv2", "conv3", "x4", "up3b", "up2b", "up1b"):
                a = col[k].detach()
                out[f"act_{k}_stats"] = np.array([a.mean().item(), a.std().item(), a.abs().max().item()])
        for name, p in ref.named_parameters():
            if p.grad is None:
                continue
            gflat = p.grad.flatten()
            out[f"grad_{tag}_{name}_norm"] = np.array([gflat.norm().item()])
            out[f"grad_{tag}_{name}_head"] = gflat[:64].numpy().copy()
    print("train forward + 36 live gradients: oracle == reference bit for bit (emb.weight: no grad)")

    out["sd_checksum"] = np.array([[v.double().sum().item(), v.double().abs().sum().item()]
                                   for v in sd.values() if v.is_floating_point()])
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "cunet_b2_h32_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
