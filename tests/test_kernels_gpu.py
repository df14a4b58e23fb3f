"""Per-kernel parity: every C-ABI entry point against a plain PyTorch fp32 restatement of the same
reference op (TF32 off), on bf16-representable inputs so the only differences are accumulation
order and the bf16 rounding of outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(t):
    return t.to(torch.bfloat16).float()


def nhwc(t):  # NCHW fp32 -> NHWC bf16
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t):  # NHWC bf16 -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


CONV_CASES = [
    # B, H, W, c0, c1, cout
    (1, 32, 32, 64, 0, 64),
    (2, 16, 16, 64, 0, 128),
    (3, 24, 40, 128, 0, 128),
    (2, 8, 8, 256, 0, 512),
    (1, 8, 16, 512, 0, 512),
    (2, 16, 16, 512, 256, 256),
    (2, 32, 32, 128, 64, 64),
    (1, 32, 32, 256, 128, 128),
    (4, 128, 128, 64, 0, 64),
    (1, 4, 4, 64, 0, 64),
]


@pytest.mark.parametrize("B,H,W,c0,c1,cout", CONV_CASES)
def test_conv3x3_fprop(cuda, B, H, W, c0, c1, cout):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H + c0 + cout)
    cin = c0 + c1
    x = bf(torch.randn(B, cin, H, W, generator=g)).to(cuda)
    w = bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b = torch.randn(cout, generator=g).to(cuda)
    wf, wd = K.pack_conv3x3_weights(w)
    s0 = nhwc(x[:, :c0])
    s1 = nhwc(x[:, c0:]) if c1 else None
    y = K.conv3x3(s0, s1, wf, b, True, None, cout)
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    assert rel(nchw(y), ref) < 6e-3
    # no bias / no relu
    y2 = K.conv3x3(s0, s1, wf, None, False, None, cout)
    ref2 = F.conv2d(x, w, None, padding=1)
    assert rel(nchw(y2), ref2) < 6e-3


@pytest.mark.parametrize("B,H,W,c0,c1,cout", CONV_CASES)
def test_conv3x3_dgrad(cuda, B, H, W, c0, c1, cout):
    """Data gradient = the same kernel on dY with the flipped/transposed pack, ReLU mask fused;
    two-source layers produce the two channel slices with row-sliced weights."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(7 + B + H + c0 + cout)
    cin = c0 + c1
    w = bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cout ** 0.5)).to(cuda)
    dy = bf(torch.randn(B, cout, H, W, generator=g)).to(cuda)
    below = bf(torch.randn(B, cin, H, W, generator=g)).clamp_min(0).to(cuda)  # post-ReLU input
    wf, wd = K.pack_conv3x3_weights(w)
    ref = F.conv_transpose2d(dy, w, padding=1)  # == conv2d input gradient
    dyh = nhwc(dy)
    if c1 == 0:
        gx = K.conv3x3(dyh, None, wd, None, False, nhwc(below), cin)
        assert rel(nchw(gx), ref * (below > 0)) < 6e-3
    else:
        g0 = K.conv3x3(dyh, None, wd[:c0], None, False, None, c0)
        g1 = K.conv3x3(dyh, None, wd[c0:], None, False, None, c1)
        assert rel(nchw(g0), ref[:, :c0]) < 6e-3
        assert rel(nchw(g1), ref[:, c0:]) < 6e-3


@pytest.mark.parametrize("B,H,W,c0,c1,cout", CONV_CASES)
def test_conv3x3_wgrad(cuda, B, H, W, c0, c1, cout):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(11 + B + H + c0 + cout)
    cin = c0 + c1
    x = bf(torch.randn(B, cin, H, W, generator=g)).to(cuda)
    dy = bf(torch.randn(B, cout, H, W, generator=g)).to(cuda)
    s0 = nhwc(x[:, :c0])
    s1 = nhwc(x[:, c0:]) if c1 else None
    dw, db = K.conv3x3_wgrad(s0, s1, nhwc(dy))
    ref_w = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dy, padding=1)
    assert rel(dw, ref_w) < 2e-3
    assert rel(db, dy.sum(dim=(0, 2, 3))) < 1e-4
    dw2, _ = K.conv3x3_wgrad(s0, s1, nhwc(dy))
    assert torch.equal(dw, dw2), "wgrad must be deterministic"


S2_CASES = [
    # B, H, W, cin, cout   (discriminator trunk shapes, scaled down, plus ragged / odd sizes)
    (2, 32, 32, 64, 128),
    (2, 16, 16, 128, 256),
    (3, 8, 8, 256, 512),
    (1, 24, 40, 64, 64),
    (2, 18, 30, 128, 128),
    (1, 17, 23, 64, 128),
    (4, 64, 64, 64, 128),
]


@pytest.mark.parametrize("B,H,W,cin,cout", S2_CASES)
def test_conv3x3_s2(cuda, B, H, W, cin, cout):
    """Stride-2 convolution (nets.py:30-31 + LeakyReLU :32): forward, data gradient, weight and
    bias gradient against PyTorch fp32 on bf16-representable inputs."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(7 * B + H + cin + cout)
    x = bf(torch.randn(B, cin, H, W, generator=g)).to(cuda)
    w = bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b = torch.randn(cout, generator=g).to(cuda)
    wf, wd = K.pack_conv3x3_weights(w)
    y = K.conv3x3_s2(nhwc(x), wf, b, 0.2, cout)
    ref = F.leaky_relu(F.conv2d(x, w, b, stride=2, padding=1), 0.2)
    assert nchw(y).shape == ref.shape
    assert rel(nchw(y), ref) < 6e-3
    y1 = K.conv3x3_s2(nhwc(x), wf, None, 1.0, cout)  # no bias, no activation
    ref1 = F.conv2d(x, w, None, stride=2, padding=1)
    assert rel(nchw(y1), ref1) < 6e-3
    Ho, Wo = ref.shape[2], ref.shape[3]
    dy = bf(torch.randn(B, cout, Ho, Wo, generator=g)).to(cuda)
    dx = K.conv3x3_s2_dgrad(nhwc(dy), wd, cin, H, W)
    ref_dx = torch.nn.grad.conv2d_input((B, cin, H, W), w, dy, stride=2, padding=1)
    assert rel(nchw(dx), ref_dx) < 6e-3
    dw, db = K.conv3x3_s2_wgrad(nhwc(x), nhwc(dy))
    ref_w = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dy, stride=2, padding=1)
    assert rel(dw, ref_w) < 2e-3
    assert rel(db, dy.sum(dim=(0, 2, 3))) < 1e-4
    dw2, _ = K.conv3x3_s2_wgrad(nhwc(x), nhwc(dy), want_bias=False)
    assert torch.equal(dw, dw2), "wgrad must be deterministic"


def test_disc_block(cuda):
    """One discriminator block (nets.py:26-33) through _DiscBlock against PyTorch fp32 autograd."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W, cin, cout = 2, 32, 32, 64, 128
    x = bf(torch.randn(B, cin, H, W, generator=g)).to(cuda)
    w0 = bf(torch.randn(cin, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b0 = (torch.randn(cin, generator=g) * 0.1).to(cuda)
    w1 = bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b1 = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    gy = bf(torch.randn(B, cout, H // 2, W // 2, generator=g)).to(cuda)
    leaves = [t.clone().requires_grad_(True) for t in (x, w0, b0, w1, b1)]
    # the block stores its intermediate activation in bf16; the fp32 reference rounds at the same
    # point (straight-through), otherwise ~0.1 % of the LeakyReLU masks differ (pre-activations
    # within the rounding error of zero) and the gradients differ by 3e-2 rel-L2 for that reason alone
    h_ref = F.conv2d(leaves[0], leaves[1], leaves[2], padding=1)
    h_ref = h_ref + (bf(h_ref) - h_ref).detach()
    ref = F.leaky_relu(F.conv2d(h_ref, leaves[3], leaves[4], stride=2, padding=1), 0.2)
    ref.backward(gy)
    mine = [t.clone().requires_grad_(True) for t in (w0, b0, w1, b1)]
    xin = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = K.disc_block(xin, *mine, 0.2)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert rel(out, ref) < 8e-3
    out.backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    # the gradient of the intermediate activation is stored in bf16 as well: 1e-2 rel-L2
    errs = {name: rel(a.grad, r.grad) for name, a, r in zip(("x", "w0", "b0", "w1", "b1"), [xin] + mine, leaves)}
    assert max(errs.values()) < 1e-2, errs


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 24, 40), (3, 64, 64), (2, 20, 18), (1, 256, 256)])
def test_conv_first(cuda, B, H, W):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(H)
    x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(cuda)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(cuda)
    b = (torch.randn(64, generator=g) * 0.1).to(cuda)
    y = K.conv_first(x, w, b)
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    assert (nchw(y) - ref).abs().max().item() < 2e-2
    assert rel(nchw(y), ref) < 4e-3
    dy = bf(torch.randn(B, 64, H, W, generator=g)).to(cuda)
    dw, db = K.conv_first_wgrad(x, nhwc(dy))
    ref_w = torch.nn.grad.conv2d_weight(x, (64, 3, 3, 3), dy, padding=1)
    assert rel(dw, ref_w) < 1e-4
    assert rel(db, dy.sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 32, 32, 64, 64), (1, 24, 40, 64, 128), (3, 72, 16, 128, 128),
                                            (1, 256, 256, 64, 64)])
def test_conv3x3_pool_fused(cuda, B, H, W, cin, cout):
    """conv3x3 + ReLU + MaxPool2d(2) in one kernel (cunet.py:45-46): both outputs bit-identical to the
    separate convolution and pooling kernels."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(B + H + cin + cout)
    x = nhwc(bf(torch.randn(B, cin, H, W, generator=g)).to(cuda))
    w = bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    wf, _ = K.pack_conv3x3_weights(w, need_dgrad=False)
    full, pooled = K.conv3x3_pool(x, wf, b, cout)
    ref_full = K.conv3x3(x, None, wf, b, True, None, cout)
    assert torch.equal(full, ref_full)
    assert torch.equal(pooled, K.maxpool2(ref_full))
    assert torch.equal(nchw(pooled), F.max_pool2d(nchw(ref_full), 2))


@pytest.mark.parametrize("B,H,W,cin", [(2, 32, 32, 64), (1, 24, 40, 64), (1, 72, 16, 128)])
def test_conv3x3_last_fused(cuda, B, H, W, cin):
    """dconv_up1.2 + conv_last + tanh in one kernel (cunet.py:78-82) against PyTorch fp32, and
    against the two separate kernels."""
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(B + H + cin)
    x = bf(torch.randn(B, cin, H, W, generator=g)).to(cuda)
    w = bf(torch.randn(64, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(cuda)
    b = (torch.randn(64, generator=g) * 0.1).to(cuda)
    lw = (torch.randn(3, 64, 1, 1, generator=g) * 0.2).to(cuda)
    lb = (torch.randn(3, generator=g) * 0.1).to(cuda)
    wf, _ = K.pack_conv3x3_weights(w, need_dgrad=False)
    h, y = K.conv3x3_last(nhwc(x), wf, b, lw, lb)
    h_ref = F.relu(F.conv2d(x, w, b, padding=1))
    assert rel(nchw(h), h_ref) < 6e-3
    y_ref = torch.tanh(F.conv2d(h_ref, lw, lb))
    assert (y - y_ref).abs().max().item() < 5e-3
    h2 = K.conv3x3(nhwc(x), None, wf, b, True, None, 64)
    assert torch.equal(h, h2)
    # the stand-alone kernel reads the bf16-rounded activations (64 terms x 2^-9 relative each)
    y2 = K.conv_last_tanh(h2, lw, lb)
    assert (y - y2).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 24, 40), (3, 64, 64)])
def test_conv_last_tanh(cuda, B, H, W):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(H + 1)
    x = bf(torch.randn(B, 64, H, W, generator=g)).clamp_min(0).to(cuda)
    w = (torch.randn(3, 64, 1, 1, generator=g) * 0.2).to(cuda)
    b = (torch.randn(3, generator=g) * 0.1).to(cuda)
    y = K.conv_last_tanh(nhwc(x), w, b)
    ref = torch.tanh(F.conv2d(x, w, b))
    assert (y - ref).abs().max().item() < 1e-5
    gy = torch.randn(B, 3, H, W, generator=g).to(cuda)
    gx, dw, db = K.conv_last_tanh_bprop(gy, y, nhwc(x), w)
    t = gy * (1 - ref * ref)
    ref_gx = F.conv_transpose2d(t, w) * (x > 0)
    assert rel(nchw(gx), ref_gx) < 4e-3
    assert rel(dw, torch.nn.grad.conv2d_weight(x, (3, 64, 1, 1), t)) < 1e-4
    assert rel(db, t.sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (1, 8, 24, 256), (3, 16, 16, 128)])
def test_maxpool(cuda, B, H, W, C):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(C)
    x = bf(torch.randn(B, C, H, W, generator=g)).clamp_min(0).to(cuda)
    y = K.maxpool2(nhwc(x))
    assert torch.equal(nchw(y), F.max_pool2d(x, 2))
    g_pool = bf(torch.randn(B, C, H // 2, W // 2, generator=g)).to(cuda)
    g_skip = bf(torch.randn(B, C, H, W, generator=g)).to(cuda)
    xr = x.clone().requires_grad_(True)
    F.max_pool2d(xr, 2).backward(g_pool)
    ref = (xr.grad + g_skip) * (x > 0)
    out = K.maxpool2_bwd(nhwc(x), nhwc(g_pool), nhwc(g_skip))
    assert rel(nchw(out), ref) < 4e-3
    out2 = K.maxpool2_bwd(nhwc(x), nhwc(g_pool), None)
    assert rel(nchw(out2), xr.grad * (x > 0)) < 1e-6


def _adain_ref(x, c, lw, lb, eps, mask, p):
    B, C = x.shape[:2]
    style = F.linear(c, lw, lb).view(B, C, 4)
    flat = x.reshape(B, C, -1)
    xs = (flat.var(-1) + eps).sqrt().view(B, C, 1, 1)
    xm = flat.mean(-1).view(B, C, 1, 1)
    ys = (style.var(-1) + eps).sqrt().view(B, C, 1, 1)
    ym = style.mean(-1).view(B, C, 1, 1)
    z = (x - xm) / xs * ys + ym
    u = F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=True)
    if mask is not None:
        u = u * mask.permute(0, 3, 1, 2).float() / (1 - p)
    return u


@pytest.mark.parametrize("B,h,w,C,nc,p", [(2, 8, 8, 512, 5, 0.3), (3, 16, 24, 128, 5, 0.3),
                                          (1, 32, 32, 256, 6, 0.0), (2, 4, 4, 128, 5, 0.3),
                                          (2, 5, 7, 128, 5, 0.3)])
def test_adain_up_drop(cuda, B, h, w, C, nc, p):
    from weather_unet_b200 import _ops as K
    g = torch.Generator(device="cpu").manual_seed(C + h)
    x = bf(torch.randn(B, C, h, w, generator=g) * 0.7 + 0.3).clamp_min(0).to(cuda)
    c = torch.randn(B, nc, generator=g).to(cuda)
    lw = (torch.randn(4 * C, nc, generator=g) * 0.4).to(cuda)
    lb = (torch.randn(4 * C, generator=g) * 0.4).to(cuda)
    mask = (torch.rand(B, 2 * h, 2 * w, C, generator=g) >= p).to(torch.uint8).to(cuda) if p > 0 else None
    u, st = K.adain_up_drop(nhwc(x), c, lw, lb, 1e-5, p, 0, mask)
    lwr, lbr = lw.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = _adain_ref(xr, c, lwr, lbr, 1e-5, mask, p)
    assert rel(nchw(u), ref) < 5e-3
    flat = x.reshape(B, C, -1)
    assert rel(st.mean, flat.mean(-1)) < 1e-5
    assert rel(st.rstd, 1 / (flat.var(-1) + 1e-5).sqrt()) < 1e-4
    gu = bf(torch.randn(B, C, 2 * h, 2 * w, generator=g)).to(cuda)
    ref.backward(gu)
    gx, dlw, dlb = K.adain_up_drop_bwd(nhwc(gu), nhwc(x), c, lw, lb, st)
    assert rel(nchw(gx), xr.grad * (x > 0)) < 8e-3
    assert rel(dlw, lwr.grad) < 2e-3
    assert rel(dlb, lbr.grad) < 2e-3


def test_dropout_philox_rate(cuda):
    """Without an injected mask the keep decisions come from Philox: rate ~ 1-p, scaling 1/(1-p),
    the same seed reproduces the mask, another seed does not."""
    from weather_unet_b200 import _ops as K
    B, h, w, C, nc = 2, 16, 16, 128, 5
    x = torch.ones(B, h, w, C, dtype=torch.bfloat16, device=cuda)
    x[:, ::2] = 3.0  # non-constant so the variance is not degenerate
    c = torch.zeros(B, nc, device=cuda)
    lw = torch.zeros(4 * C, nc, device=cuda)
    lb = torch.tensor([0.0, 2.0, 0.0, 2.0], device=cuda).repeat(C)  # y_mean 1, y_std > 0
    u0, _ = K.adain_up_drop(x, c, lw, lb, 1e-5, 0.0, 5, None)
    u1, st1 = K.adain_up_drop(x, c, lw, lb, 1e-5, 0.3, 5, None)
    u2, _ = K.adain_up_drop(x, c, lw, lb, 1e-5, 0.3, 5, None)
    u3, _ = K.adain_up_drop(x, c, lw, lb, 1e-5, 0.3, 6, None)
    assert torch.equal(u1, u2) and not torch.equal(u1, u3)
    kept = u1 != 0
    nz = u0 != 0
    rate = (kept & nz).sum().item() / nz.sum().item()
    assert abs(rate - 0.7) < 0.01
    sel = kept & nz
    assert rel(u1[sel].float(), u0[sel].float() / 0.7) < 5e-3
    # the keep byte the backward pass reads (bit j = channel 8v + j) is the decision the forward applied
    bits = ((st1.bits.unsqueeze(-1).int() >> torch.arange(8, device=cuda)) & 1).bool().reshape(u1.shape)
    assert torch.equal(bits[nz], kept[nz])
    # and the backward honours it: dropped positions pass no gradient
    gu = torch.ones_like(u1)
    gx_keep, _, _ = K.adain_up_drop_bwd(gu, x, c, lw, lb, st1)
    st1.bits.zero_()
    gx_none, _, _ = K.adain_up_drop_bwd(gu, x, c, lw, lb, st1)
    assert gx_none.float().abs().max().item() == 0.0 and gx_keep.float().abs().max().item() > 0.0


def test_layout_roundtrip(cuda):
    from weather_unet_b200 import _ops as K
    x = bf(torch.randn(2, 70, 9, 13)).to(cuda)
    y = K.nchw_to_nhwc(x)
    assert torch.equal(y, nhwc(x))
    assert torch.equal(K.nhwc_to_nchw(y), x)


def test_errors(cuda):
    from weather_unet_b200 import _ops as K
    from weather_unet_b200._lib import WuError
    x = torch.zeros(1, 8, 8, 48, dtype=torch.bfloat16, device=cuda)
    w = torch.zeros(64, 9 * 48, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(WuError, match="multiple of 64"):
        K.conv3x3(x, None, w, None, True, None, 64)
    with pytest.raises(WuError):
        K.maxpool2(torch.zeros(1, 7, 8, 64, dtype=torch.bfloat16, device=cuda))


def test_standalone_blocks(cuda):
    """r_double_conv and AdaIN used on their own (forward only) match the oracle's functions."""
    from oracle import cunet_oracle as orc
    from weather_unet_b200.nets import r_double_conv
    from weather_unet_b200.utils import AdaIN
    torch.manual_seed(4)
    blk = r_double_conv(64, 128).to(cuda)
    x = bf(torch.randn(2, 64, 16, 24)).to(cuda)
    sd = {f"b.{k}": v for k, v in blk.state_dict().items()}
    with torch.no_grad():
        y = blk(x)
        ref = orc.double_conv(sd, "b", x)
    assert y.shape == ref.shape and rel(y, ref) < 8e-3
    first = r_double_conv(3, 64).to(cuda)
    xi = torch.rand(2, 3, 16, 24, device=cuda) * 2 - 1
    with torch.no_grad():
        assert rel(first(xi), orc.double_conv({f"b.{k}": v for k, v in first.state_dict().items()}, "b", xi)) < 8e-3
    ad = AdaIN(128, num_classes=5).to(cuda)
    xa = bf(torch.randn(2, 128, 8, 8)).to(cuda)
    c = torch.randn(2, 5, device=cuda)
    with torch.no_grad():
        ya = ad(xa, c)
        ra = orc.adain({f"a.{k}": v for k, v in ad.state_dict().items()}, "a", xa, c)
    assert rel(ya, ra) < 6e-3
    with pytest.raises(RuntimeError, match="forward-only"):
        blk(x.requires_grad_(True))


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 32, 32, 256, 512), (3, 24, 40, 128, 256), (2, 48, 40, 128, 128),
                                            (1, 40, 24, 64, 128), (2, 16, 16, 512, 512)])
def test_conv3x3_fused_adain_stats(cuda, B, H, W, cin, cout):
    """wu_conv3x3_fprop_stats: the convolution output is bit-identical to wu_conv3x3_fprop, and the
    (sum, sum of squares) it emits fold to the statistics wu_adain_stats computes from the stored
    tensor (utils.py:34-39), including ragged sizes whose tiles hang over the image edge."""
    from weather_unet_b200 import _ops as K
    from weather_unet_b200._lib import call, ptr, query, stream
    g = torch.Generator().manual_seed(B * H + cout)
    x = torch.randn(B, H, W, cin, generator=g).to(cuda).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(cuda)
    bias = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    wf, _ = K.pack_conv3x3_weights(w)
    ref = K.conv3x3(x, None, wf, bias, True, None, cout)
    dst, sums = K.conv3x3_stats(x, None, wf, bias, cout)
    assert sums is not None and sums.shape[0] == B and sums.shape[2:] == (cout, 2)
    assert torch.equal(dst, ref)
    tot = sums.double().sum(dim=1)  # (B, cout, 2)
    xf = dst.float().double().reshape(B, H * W, cout)
    assert torch.allclose(tot[..., 0], xf.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(tot[..., 1], (xf * xf).sum(1), rtol=1e-5, atol=1e-3)
    # and through the AdaIN site: same scale / shift as the separate statistics pass
    cond = torch.randn(B, 5, generator=g).to(cuda)
    lw, lb = (torch.randn(4 * cout, 5, generator=g) * 0.3).to(cuda), torch.zeros(4 * cout, device=cuda)
    u1, st1 = K.adain_up_drop(dst, cond, lw, lb, 1e-5, 0.0, 0, None)
    u2, st2 = K.adain_up_drop(dst, cond, lw, lb, 1e-5, 0.0, 0, None, stats=sums)
    assert torch.allclose(st1.mean, st2.mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(st1.rstd, st2.rstd, rtol=1e-4)
    assert ((u1.float() - u2.float()).abs().max() <= 2e-2 * u1.float().abs().max()).item()
