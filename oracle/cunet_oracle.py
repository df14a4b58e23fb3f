"""ORACLE (test infrastructure, not product code): a functional PyTorch-fp32 restatement of the
reference generator `Conditional_UNet.forward` (reference cunet.py:43-82) with `r_double_conv`
(nets.py:18-24) and `AdaIN` (utils.py:34-51), operating on a plain state_dict.

Parity status: PINNED against the unmodified reference imported from /root/reference in the build
container (oracle/pin_against_reference.py; fixtures in tests/golden/).  The reference ships no
golden vectors or tests of its own, so "pinned" means: bit-identical outputs and
gradients to the live reference modules on CPU for the committed seeds, in eval mode and in train
mode with the same RNG stream.

Every function works on any device/dtype torch supports; tests use fp32 (TF32 disabled on GPU).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5        # AdaIN eps (utils.py:27, passed to c_norm at utils.py:47-48)
P_DROP = 0.3      # nn.Dropout(p=0.3) (cunet.py:28)
BLOCKS = ("dconv_down1", "dconv_down2", "dconv_down3", "dconv_down4", "dconv_up3", "dconv_up2",
          "dconv_up1")


class _RoundBF16(torch.autograd.Function):
    """bf16 storage emulation: value and incoming gradient are rounded to bf16 (kept as fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGradBF16(torch.autograd.Function):
    """Identity forward, bf16-rounded gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundValueBF16(torch.autograd.Function):
    """bf16-rounded value, gradient passed through unrounded."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _flags(q):
    """emulate_bf16 may be a bool (all / nothing) or a subset of {"w", "a", "g"}: round the
    convolution Weights, the stored Activations, the activation Gradients (tools/
    grad_error_attribution.py separates their contributions)."""
    if q is True:
        return frozenset("wag")
    if not q:
        return frozenset()
    return frozenset(q)


def _q(x, on):
    f = _flags(on)
    if "a" in f and "g" in f:
        return _RoundBF16.apply(x)
    if "a" in f:
        return _RoundValueBF16.apply(x)
    if "g" in f:
        return _RoundGradBF16.apply(x)
    return x


def _qw(w, on):
    """bf16-rounded weight in the forward/data-gradient, fp32 master weight receives the gradient."""
    if on is True or (on and "w" in _flags(on)):
        return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach()
    return w


def double_conv(sd, name, x, q=False, q_first=True, tap=None):
    """nets.py:18-24: conv3x3(pad 1) -> ReLU -> conv3x3(pad 1) -> ReLU.
    q: emulate the CUDA path's storage precision (bf16 weights and activations, fp32 accumulate);
    q_first=False keeps the first convolution's weights in fp32 (the K=27 image layer does).
    tap: optional callable(name_suffix, tensor) -> tensor applied to both activations."""
    tap = tap or (lambda k, v: v)
    x = _q(F.relu(F.conv2d(x, _qw(sd[f"{name}.0.weight"], q if q_first else False),
                           sd[f"{name}.0.bias"], padding=1)), q)
    x = tap(f"{name}.a", x)
    x = _q(F.relu(F.conv2d(x, _qw(sd[f"{name}.2.weight"], q), sd[f"{name}.2.bias"], padding=1)), q)
    return tap(f"{name}.b", x)


def adain(sd, name, x, c, eps=EPS):
    """utils.py:41-51.  style = l1(c) viewed (B, C, 4); both statistics use the UNBIASED variance
    (torch default, utils.py:36) with eps added to the variance before the square root."""
    B, C = x.shape[:2]
    style = F.linear(c, sd[f"{name}.l1.weight"], sd[f"{name}.l1.bias"]).view(B, C, 4)
    flat = x.reshape(B, C, -1)
    x_std = (flat.var(dim=-1) + eps).sqrt().view(B, C, 1, 1)
    x_mean = flat.mean(dim=-1).view(B, C, 1, 1)
    y_std = (style.var(dim=-1) + eps).sqrt().view(B, C, 1, 1)
    y_mean = style.mean(dim=-1).view(B, C, 1, 1)
    return (x - x_mean) / x_std * y_std + y_mean


def upsample(x):
    """cunet.py:26: bilinear x2, align_corners=True."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def dropout(x, mask, train, p=P_DROP):
    """cunet.py:28.  mask: None -> draw from torch's RNG exactly like nn.Dropout's CPU path
    (bernoulli_(1-p) noise of x's shape, scaled by 1/(1-p)); else a uint8 NHWC keep mask."""
    if not train or p == 0.0:
        return x, None
    if mask is None:
        keep = torch.empty_like(x).bernoulli_(1 - p)
    else:
        keep = mask.permute(0, 3, 1, 2).to(x.dtype)
    return x * (keep / (1 - p)), keep


# name of each activation in the CUDA path's bookkeeping (weather-unet_b200/_generator.py)
ACT_NAMES = {"dconv_down1.a": "a1", "dconv_down1.b": "conv1", "dconv_down2.a": "d2a",
             "dconv_down2.b": "conv2", "dconv_down3.a": "d3a", "dconv_down3.b": "conv3",
             "dconv_down4.a": "d4a", "dconv_down4.b": "x4", "dconv_up3.a": "up3a",
             "dconv_up3.b": "up3b", "dconv_up2.a": "up2a", "dconv_up2.b": "up2b",
             "dconv_up1.a": "up1a", "dconv_up1.b": "up1b"}


def forward(sd, x, c, train=False, masks=None, p=P_DROP, collect=None, emulate_bf16=False,
            override=None):
    """cunet.py:43-82.  `masks`: optional 3 uint8 NHWC keep masks (sites adain3, adain2, adain1).
    `collect`: optional dict filled with named intermediate activations (NCHW) and drawn masks.
    `emulate_bf16`: NOT the reference's arithmetic — the same dataflow with values rounded to bf16
    wherever the CUDA path stores bf16 (conv weights, activations, activation gradients).
    `override`: dict name -> NCHW tensor; each named activation takes the given VALUE while keeping
    its local derivative ("teacher forcing"), so that ReLU masks / pooling arg-maxes of a backward
    pass agree with the run the values came from and gradient parity can be checked tightly."""
    q = emulate_bf16
    masks = masks or (None, None, None)

    def tap(k, v):
        k = ACT_NAMES.get(k, k)
        if override is not None and k in override:
            v = v + (override[k].to(v.dtype) - v).detach()
        if collect is not None:
            collect[k] = v
        return v

    conv1 = double_conv(sd, "dconv_down1", x, q, q_first=False, tap=tap)
    conv2 = double_conv(sd, "dconv_down2", F.max_pool2d(conv1, 2), q, tap=tap)
    conv3 = double_conv(sd, "dconv_down3", F.max_pool2d(conv2, 2), q, tap=tap)
    h = double_conv(sd, "dconv_down4", F.max_pool2d(conv3, 2), q, tap=tap)
    for i, (ad, up, skip) in enumerate((("adain3", "dconv_up3", conv3),
                                        ("adain2", "dconv_up2", conv2),
                                        ("adain1", "dconv_up1", conv1))):
        z = adain(sd, ad, h, c)
        h = upsample(_RoundGradBF16.apply(z) if "g" in _flags(q) else z)
        h, m = dropout(h, masks[i], train, p)
        h = tap(f"u{3 - i}", _q(h, q))
        if collect is not None:
            collect[f"mask{3 - i}"] = m
        h = double_conv(sd, up, torch.cat([h, skip], dim=1), q, tap=tap)  # [x, skip] (cunet.py:62)
    out = F.conv2d(h, sd["conv_last.weight"], sd["conv_last.bias"])
    return torch.tanh(out)


def forward_backward(sd, x, c, masks, gy, train=True, p=P_DROP, emulate_bf16=False):
    """Forward + backward of sum(y * gy).  Returns (y, {param name: grad})."""
    leaf = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    y = forward(leaf, x, c, train=train, masks=masks, p=p, emulate_bf16=emulate_bf16)
    (y * gy).sum().backward()
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    return y.detach(), grads


def make_dropout_masks(B, H, W, p=P_DROP, seed=0, device="cpu"):
    """Three uint8 NHWC keep masks for the dropout sites at H/4 (512 ch), H/2 (256), H (128)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(B, H // 4, W // 4, 512), (B, H // 2, W // 2, 256), (B, H, W, 128)]
    return tuple((torch.rand(s, generator=g) >= p).to(torch.uint8).to(device) for s in shapes)


def conv_flops(H, W, B=1):
    """2*MAC of the 15 convolutions of one forward pass."""
    total, res = 0, (H, W)
    plan = [(3, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4),
            (256, 512, 8), (512, 512, 8), (768, 256, 4), (256, 256, 4), (384, 128, 2), (128, 128, 2),
            (192, 64, 1), (64, 64, 1)]
    for cin, cout, d in plan:
        total += 2 * 9 * cin * cout * (res[0] // d) * (res[1] // d)
    total += 2 * 64 * 3 * H * W
    return total * B
