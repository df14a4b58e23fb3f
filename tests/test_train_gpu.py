"""Training-step parity on the GPU: the fused bias+LeakyReLU kernels, the discriminator fast path,
and a fixed-seed loss curve of GDTrainStep against the curve recorded from the reference modules
(tests/golden/train_b2_h32_seed0.npz), with the same dropout masks injected."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_b2_h32_seed0.npz")


@pytest.mark.parametrize("B,C,H,W,slope", [(2, 64, 16, 16, 0.2), (3, 128, 8, 24, 1.0), (1, 512, 4, 4, 0.2)])
def test_bias_act(cuda, B, C, H, W, slope):
    from weather_unet_b200 import _ops as K
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(C, generator=g).to(cuda)
    gy = torch.randn(B, C, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xr = x.float().clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.leaky_relu(xr + br.view(1, -1, 1, 1), slope)
    ref.backward(gy.float())
    xin = x.clone().requires_grad_(True)
    bin_ = b.clone().requires_grad_(True)
    y = K.bias_act(xin * 1.0, bin_, slope)  # xin * 1.0: a non-leaf the op may overwrite
    assert ((y.float() - ref).norm() / ref.norm()).item() < 4e-3
    y.backward(gy)
    assert ((xin.grad.float() - xr.grad).norm() / xr.grad.norm()).item() < 4e-3
    assert ((bin_.grad - br.grad).norm() / br.grad.norm()).item() < 2e-3


def test_fused_spectral_norm(cuda):
    """wu_sn_forward / wu_sn_backward against torch.nn.utils.spectral_norm's hook (the reference's
    mechanism): W / sigma, the updated u / v buffers over several training forwards, eval mode, and
    the gradient w.r.t. weight_orig."""
    import torch.nn as nn
    from weather_unet_b200._spectral import FusedSpectralNorm

    def make():
        torch.manual_seed(3)
        return [nn.utils.spectral_norm(nn.Conv2d(64, 128, 3, padding=1)).to(cuda),
                nn.utils.spectral_norm(nn.Conv2d(3, 3, 3, padding=1)).to(cuda),
                nn.utils.spectral_norm(nn.Conv2d(256, 512, 3, padding=1, stride=2)).to(cuda),
                nn.utils.spectral_norm(nn.Linear(512, 1)).to(cuda),
                nn.utils.spectral_norm(nn.Linear(5, 512)).to(cuda)]

    mine, ref = make(), make()
    sn = FusedSpectralNorm(mine)
    assert sn.supported()
    gen = torch.Generator().manual_seed(4)
    for it in range(3):
        training = it < 2
        for m in mine + ref:
            m.train(training)
        ws, packed = sn(training)
        for m in ref:  # the hook: power iteration (training) + weight = weight_orig / sigma
            for h in m._forward_pre_hooks.values():
                h(m, None)
        gs = [torch.randn(w.shape, generator=gen).to(cuda) for w in ws]
        torch.autograd.backward(ws, gs)
        torch.autograd.backward([m.weight for m in ref], gs)
        from weather_unet_b200 import _ops as K
        for i, (w, m, r) in enumerate(zip(ws, mine, ref)):
            if packed[i] is None:
                assert torch.allclose(w, r.weight, rtol=2e-5, atol=1e-7), (it, w.shape)
            else:  # 3x3 weights of the tcgen05 kernels: only the packed bf16 layouts are written
                wf, wd = K.pack_conv3x3_weights(r.weight.detach().contiguous())
                for a, b in zip(packed[i], (wf, wd)):
                    assert (a.float() - b.float()).abs().max().item() <= 1e-2 * b.float().abs().max().item()
                    assert (a != b).float().mean().item() < 1e-3  # same values up to rare 1-ulp bf16 ties
            assert torch.allclose(m.weight_u, r.weight_u, rtol=1e-4, atol=1e-6)
            assert torch.allclose(m.weight_v, r.weight_v, rtol=1e-4, atol=1e-6)
            e = ((m.weight_orig.grad - r.weight_orig.grad).norm() / r.weight_orig.grad.norm()).item()
            assert e < 1e-5, (it, w.shape, e)
            m.weight_orig.grad = None
            r.weight_orig.grad = None


def test_disc_fast_path(cuda):
    """SNDisc under bf16 autocast on an fp32 NCHW image runs entirely on the sm_100a kernels (fused
    spectral norm, stem, tcgen05 trunk).  Output, every parameter gradient and the spectral-norm
    buffers are checked against the fp32 ORACLE (oracle/train_oracle.disc_forward, pinned bit-exact
    to the reference's disc.py:27-38), with stock PyTorch bf16 autocast of the same modules run
    alongside as the error band a bf16 discriminator has against fp32."""
    from oracle import train_oracle as T
    from weather_unet_b200.disc import SNDisc
    from weather_unet_b200 import _ops as K
    torch.manual_seed(100)
    d1 = SNDisc(5).to(cuda).train()
    torch.manual_seed(100)
    d2 = SNDisc(5).to(cuda).train().to(memory_format=torch.channels_last)
    sd = {k: v.detach().clone() for k, v in d1.state_dict().items()}
    for k in sd:
        if k.endswith("_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    x = torch.rand(4, 3, 64, 64, device=cuda) * 2 - 1
    c = torch.eye(5, device=cuda)[:4]
    n0 = K.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res = d1(x, c)
    o1 = res[0].float()
    # 4 spectral-norm launches (which also write the packed bf16 weights), 2 stem, 2 per trunk block
    assert K.launch_count() - n0 >= 4 + 2 + 3 * 2, "kernel path not taken"
    assert [tuple(f.shape) for f in res[1:]] == [(4, 64, 32, 32), (4, 128, 16, 16), (4, 256, 8, 8), (4, 512, 4, 4)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o2 = d2(x.contiguous(memory_format=torch.channels_last), c)[0].float()  # plain modules (cuDNN bf16)
    ref = T.disc_forward(sd, x, c, train=True)  # fp32, TF32 off (conftest)
    o_ref = ref[0]
    scale = o_ref.abs().max().item()
    e1 = (o1 - o_ref).abs().max().item() / scale
    e2 = (o2 - o_ref).abs().max().item() / scale
    print(f"output vs fp32 oracle: kernels {e1:.2e}, cuDNN autocast {e2:.2e}")
    assert e1 < 2e-2 and e1 < 2 * e2 + 5e-3
    for f1, fr in zip(res[1:], ref[1:]):  # the feature maps the reference also returns
        assert ((f1.float() - fr).norm() / fr.norm()).item() < 1.5e-2
    for n, b1 in d1.named_buffers():  # power iteration ran in fp32 here, like the oracle's
        assert torch.allclose(b1, sd[n], rtol=1e-4, atol=1e-6), n
    o1.sum().backward()
    o2.sum().backward()
    o_ref.sum().backward()
    rows = []
    for (n, p), (_, q) in zip(d1.named_parameters(), d2.named_parameters()):
        g = sd[n].grad
        r1 = ((p.grad.float() - g).norm() / (g.norm() + 1e-12)).item()
        r2 = ((q.grad.float() - g).norm() / (g.norm() + 1e-12)).item()
        cos = (p.grad.float().flatten() @ g.flatten() / (p.grad.float().norm() * g.norm() + 1e-30)).item()
        print(f"   {n:28s} rel-L2 vs fp32 oracle: kernels {r1:.3e}  cuDNN autocast {r2:.3e}  cos {cos:.5f}")
        rows.append((n, r1, r2, cos))
    # SURVEY §8c proposes rel-L2 <= 3e-2 and cosine >= 0.999 vs fp32.  That holds for every tensor but
    # the stem (the deepest gradients: eight bf16 layers of backward behind them), where stock bf16
    # autocast itself is 7.6e-2 off fp32; there the bound is the autocast band.
    for n, r1, r2, cos in rows:
        assert r1 < 3e-2 or r1 < 1.25 * r2 + 5e-3, f"{n}: kernels {r1} vs fp32, autocast band {r2}"
        assert cos > 0.998, f"{n}: cos {cos}"


@pytest.mark.parametrize("d_kernels", [False, True])
def test_loss_curve_against_reference_golden(cuda, d_kernels):
    """d_kernels=False: fp32 PyTorch discriminator (isolates the generator kernels);
    d_kernels=True: the whole iteration on the library (bf16 discriminator kernels, fused spectral
    norm, fused L1, multi-tensor Adam) against the same fp32 reference curve."""
    from weather_unet_b200 import Conditional_UNet
    from weather_unet_b200.disc import SNDisc
    from weather_unet_b200.train_step import GDTrainStep
    z = np.load(GOLD)
    torch.manual_seed(0)
    G = Conditional_UNet(5).to(cuda).train()
    torch.manual_seed(100)
    D = SNDisc(5).to(cuda).train()
    step = GDTrainStep(G, D, lr=float(z["lr"][0]), d_autocast=d_kernels)
    x, cr, ct = (torch.from_numpy(z[k]).to(cuda) for k in ("images", "c_real", "c_target"))
    keys = [str(k) for k in z["keys"]]
    B, H = x.shape[0], x.shape[2]
    shapes = [(B, H // 4, H // 4, 512), (B, H // 2, H // 2, 256), (B, H, H, 128)]

    def masks(i):
        out = []
        for j, s in enumerate(shapes):
            bits = np.unpackbits(z[f"mask_{i}_{j}"])[:int(np.prod(s))]
            out.append(torch.from_numpy(bits.reshape(s).astype(np.uint8)).to(cuda))
        return tuple(out)

    curve = []
    for it in range(z["curve"].shape[0]):
        out = step.step(x, cr, ct, masks_d=masks(2 * it), masks_g=masks(2 * it + 1))
        curve.append([float(out[k]) for k in keys])
    curve = np.array(curve)
    print(np.array2string(curve, precision=4), "\n", np.array2string(z["curve"], precision=4))
    # fixed seed, injected masks: per-step |dloss| / max(|loss|, 1) <= 5e-2 on the
    # large terms; g_loss_adv / d_loss sit near zero and chaotic GAN dynamics at lr 1e-3 amplify
    # bf16 differences after a few iterations, so they are held to the first 3 iterations
    big = [keys.index(k) for k in ("g_loss", "loss_con", "g_loss_l1")]
    ref = z["curve"]
    err = np.abs(curve - ref) / np.maximum(np.abs(ref), 1.0)
    print("relative error per iteration / term:\n", np.array2string(err, precision=3))
    assert err[:, big].max() < 5e-2, err
    assert err[:3].max() < 1.5e-1, err


def test_fused_adam_matches_torch(cuda):
    """wu_adam_multi == torch.optim.Adam (the reference's optimiser settings) over several steps,
    including a parameter that never receives a gradient and a channels_last parameter."""
    from weather_unet_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 3, 3, 3), (64,), (256, 768, 3, 3), (2048, 5), (5, 5), (100001,)]
    ps_a = [torch.randn(s, device=cuda).requires_grad_(True) for s in shapes]
    ps_a[0] = ps_a[0].detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ps_b = [p.detach().clone(memory_format=torch.preserve_format).requires_grad_(True) for p in ps_a]
    lr = 1e-3
    oa = FusedAdam(ps_a, lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
    ob = torch.optim.Adam(ps_b, lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
    g = torch.Generator(device="cpu").manual_seed(1)
    for it in range(4):
        for i, (a, b) in enumerate(zip(ps_a, ps_b)):
            if i == 4:
                continue  # never gets a gradient (like adain*.emb.weight)
            gr = torch.randn(shapes[i], generator=g).to(cuda)
            a.grad = gr.clone().contiguous(memory_format=torch.channels_last) if i == 0 else gr.clone()
            b.grad = a.grad.clone(memory_format=torch.preserve_format)
        oa.step()
        ob.step()
    for a, b in zip(ps_a, ps_b):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max()
    assert torch.equal(ps_a[4], ps_b[4])
    oa2 = FusedAdam(ps_a, lr=lr, betas=(0.5, 0.9), eps=1e-6, weight_decay=0.0)
    ob2 = torch.optim.Adam(ps_b, lr=lr, betas=(0.5, 0.9), eps=1e-6, weight_decay=0.0)
    for it in range(3):
        oa2.step()
        ob2.step()
    for a, b in zip(ps_a, ps_b):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_disc_stem_kernels(cuda):
    """wu_conv3to3_* / wu_conv3to64_s2_* (discriminator stem) against PyTorch fp32 convolutions."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator().manual_seed(9)
    B, H, W = 3, 32, 48
    x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(cuda)
    w0 = (torch.randn(3, 3, 3, 3, generator=g) * 0.4).to(cuda)
    b0 = (torch.randn(3, generator=g) * 0.1).to(cuda)
    w1 = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).to(cuda)
    b1 = (torch.randn(64, generator=g) * 0.1).to(cuda)
    gy = torch.randn(B, 64, H // 2, W // 2, generator=g).to(cuda)
    leaves = [t.clone().requires_grad_(True) for t in (x, w0, b0, w1, b1)]
    ref = F.leaky_relu(F.conv2d(F.conv2d(leaves[0], leaves[1], leaves[2], padding=1), leaves[3], leaves[4],
                                stride=2, padding=1), 0.2)
    ref.backward(gy)
    mine = [t.clone().requires_grad_(True) for t in (x, w0, b0, w1, b1)]
    out = K.disc_stem(*mine, 0.2)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 4e-3
    out.backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    # the reference's gradient mask follows its own fp32 sign; ours follows the bf16 output: equal
    # except where the pre-activation is ~0, hence rel-L2 tolerances of 1e-2
    for name, a, r in zip(("x", "w0", "b0", "w1", "b1"), mine, leaves):
        e = ((a.grad - r.grad).norm() / r.grad.norm()).item()
        assert e < 1.5e-2, f"{name}: {e}"
    # no gradient requested for the image (real images, detached fakes): g_x is skipped
    mine2 = [x.clone()] + [t.clone().requires_grad_(True) for t in (w0, b0, w1, b1)]
    K.disc_stem(*mine2, 0.2).backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    assert torch.allclose(mine2[1].grad, mine[1].grad) and torch.allclose(mine2[3].grad, mine[3].grad)


def test_l1_per_sample(cuda):
    """wu_l1_per_sample_* (t_cls_train.py:255,259-266) against the PyTorch expression, value and
    gradient."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator().manual_seed(3)
    a = (torch.rand(5, 3, 24, 40, generator=g) * 2 - 1).to(cuda).requires_grad_(True)
    b = (torch.rand(5, 3, 24, 40, generator=g) * 2 - 1).to(cuda)
    w = torch.rand(5, generator=g).to(cuda)
    ref = (a - b).abs().mean(dim=(1, 2, 3))
    (ref * w).sum().backward()
    g_ref, a.grad = a.grad, None
    assert K.l1_per_sample_supported(a, b)
    d = K.l1_per_sample(a, b)
    assert torch.allclose(d, ref, rtol=1e-5, atol=1e-7)
    (d * w).sum().backward()
    assert torch.allclose(a.grad, g_ref, rtol=1e-6, atol=1e-9)
