"""Tensor-level wrappers over the C ABI (one function per entry point of include/wu_b200.h).

Every wrapper allocates its outputs / scratch with torch (the library never allocates), passes raw
device pointers plus the current CUDA stream, and returns torch tensors.  Activations are NHWC
bf16 tensors of shape (B, H, W, C).
"""
import torch

from . import _lib
from ._lib import call, on_tensor_device, ptr, query, stream

BF16 = torch.bfloat16


def _act(B, H, W, C, like):
    return torch.empty((B, H, W, C), dtype=BF16, device=like.device)


def pack_conv3x3_weights(w, need_dgrad=True):
    """fp32 [cout][cin][3][3] -> (bf16 [cout][9*cin], bf16 [cin][9*cout] or None)."""
    cout, cin = w.shape[0], w.shape[1]
    wf = torch.empty((cout, 9 * cin), dtype=BF16, device=w.device)
    wd = torch.empty((cin, 9 * cout), dtype=BF16, device=w.device) if need_dgrad else None
    call("wu_pack_conv3x3_weights", ptr(w), cout, cin, ptr(wf), ptr(wd), stream())
    return wf, wd


def pack_conv3x3_weights_into(w, wf, wd):
    """Same, into caller-held buffers (wd may be None)."""
    call("wu_pack_conv3x3_weights", ptr(w), w.shape[0], w.shape[1], ptr(wf), ptr(wd), stream())


def conv3x3(src0, src1, w_packed, bias, relu, mask, cout, src1_bcast=False):
    """3x3/s1/p1 convolution over the (virtual) channel concat [src0, src1]; see wu_conv3x3_fprop.
    src1_bcast: src1 has batch 1 and is shared by every image of src0."""
    B, H, W, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    dst = _act(B, H, W, cout, src0)
    call("wu_conv3x3_fprop_bcast", ptr(src0), c0, ptr(src1), c1, int(src1_bcast), ptr(w_packed),
         ptr(bias), int(relu), ptr(mask), ptr(dst), cout, B, H, W, stream())
    return dst


def conv3x3_stats(src0, src1, w_packed, bias, cout, src1_bcast=False):
    """conv3x3 + ReLU that also returns AdaIN's (sum, sum of squares) partials of its output
    (wu_conv3x3_fprop_stats): -> (dst, stats (B, chunks, cout, 2) fp32), or (dst, None) when the shape
    has no fused variant (the caller then runs wu_adain_stats)."""
    B, H, W, c0 = src0.shape
    chunks = query("wu_conv3x3_stats_chunks", cout, H, W)
    if chunks <= 0:
        return conv3x3(src0, src1, w_packed, bias, True, None, cout, src1_bcast=src1_bcast), None
    c1 = 0 if src1 is None else src1.shape[3]
    dst = _act(B, H, W, cout, src0)
    stats = torch.empty((B, chunks, cout, 2), dtype=torch.float32, device=src0.device)
    call("wu_conv3x3_fprop_stats", ptr(src0), c0, ptr(src1), c1, int(src1_bcast), ptr(w_packed),
         ptr(bias), ptr(dst), ptr(stats), cout, B, H, W, stream())
    return dst, stats


def conv3x3_last(src, w_packed, bias, last_w, last_b):
    """dconv_up1.2 + conv_last + tanh in one kernel (wu_conv3x3_fprop_last):
    -> (relu(conv3x3(src)) NHWC bf16 (B,H,W,64), tanh(conv1x1) fp32 NCHW (B,3,H,W))."""
    B, H, W, cin = src.shape
    dst = _act(B, H, W, 64, src)
    y = torch.empty((B, 3, H, W), dtype=torch.float32, device=src.device)
    call("wu_conv3x3_fprop_last", ptr(src), cin, ptr(w_packed), ptr(bias), ptr(dst), ptr(last_w),
         ptr(last_b), ptr(y), B, H, W, stream())
    return dst, y


def conv3x3_pool(src, w_packed, bias, cout):
    """conv3x3 + ReLU + MaxPool2d(2) in one kernel (wu_conv3x3_fprop_pool):
    -> (full-resolution output (B,H,W,cout), pooled (B,H/2,W/2,cout))."""
    B, H, W, cin = src.shape
    dst = _act(B, H, W, cout, src)
    pooled = _act(B, H // 2, W // 2, cout, src)
    call("wu_conv3x3_fprop_pool", ptr(src), cin, ptr(w_packed), ptr(bias), ptr(dst), ptr(pooled), cout,
         B, H, W, stream())
    return dst, pooled


def conv3x3_wgrad(src0, src1, dy, want_bias=True, dw=None, db=None):
    """-> (dw fp32 [cout][cin][3][3], db fp32 [cout] or None).  dw / db: optional caller-held
    destinations (e.g. views into a flat gradient bucket)."""
    B, H, W, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    cout = dy.shape[3]
    cin = c0 + c1
    nbytes = query("wu_conv3x3_wgrad_workspace_bytes", cin, cout, B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dy.device)
    if dw is None:
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dy.device)
    if want_bias and db is None:
        db = torch.empty((cout,), dtype=torch.float32, device=dy.device)
    call("wu_conv3x3_wgrad", ptr(src0), c0, ptr(src1), c1, ptr(dy), cout, B, H, W, ptr(dw), ptr(db),
         ptr(ws), nbytes, stream())
    return dw, db


def conv_first(x, w, bias):
    """x fp32 NCHW (B,3,H,W) -> relu(conv3x3) NHWC bf16 (B,H,W,64)."""
    B, _, H, W = x.shape
    dst = _act(B, H, W, 64, x)
    call("wu_conv_first_fprop", ptr(x), ptr(w), ptr(bias), ptr(dst), B, H, W, stream())
    return dst


def conv_first_wgrad(x, dy, dw=None, db=None):
    B, _, H, W = x.shape
    nbytes = query("wu_conv_first_wgrad_workspace_bytes", B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    if dw is None:
        dw = torch.empty((64, 3, 3, 3), dtype=torch.float32, device=x.device)
    if db is None:
        db = torch.empty((64,), dtype=torch.float32, device=x.device)
    call("wu_conv_first_wgrad", ptr(x), ptr(dy), ptr(dw), ptr(db), B, H, W, ptr(ws), nbytes, stream())
    return dw, db


def conv_last_tanh(x, w, bias):
    """x NHWC bf16 (B,H,W,64) -> tanh(conv1x1) fp32 NCHW (B,3,H,W)."""
    B, H, W, _ = x.shape
    y = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    call("wu_conv_last_tanh_fprop", ptr(x), ptr(w), ptr(bias), ptr(y), B, H, W, stream())
    return y


def conv_last_tanh_bprop(gy, y, x, w, dw=None, db=None):
    B, H, W, _ = x.shape
    nbytes = query("wu_conv_last_tanh_bprop_workspace_bytes", B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    gx = torch.empty_like(x)
    if dw is None:
        dw = torch.empty((3, 64, 1, 1), dtype=torch.float32, device=x.device)
    if db is None:
        db = torch.empty((3,), dtype=torch.float32, device=x.device)
    call("wu_conv_last_tanh_bprop", ptr(gy), ptr(y), ptr(x), ptr(w), ptr(gx), ptr(dw), ptr(db), B, H,
         W, ptr(ws), nbytes, stream())
    return gx, dw, db


def maxpool2(src):
    B, H, W, C = src.shape
    dst = _act(B, H // 2, W // 2, C, src)
    call("wu_maxpool2_fwd", ptr(src), ptr(dst), B, H, W, C, stream())
    return dst


def maxpool2_bwd(y, g_pool, g_skip):
    B, H, W, C = y.shape
    g = torch.empty_like(y)
    call("wu_maxpool2_bwd", ptr(y), ptr(g_pool), ptr(g_skip), ptr(g), B, H, W, C, stream())
    return g


class AdaINState:
    """Per-call statistics of one AdaIN site kept for the backward pass (all fp32 [B][C])."""
    __slots__ = ("mean", "rstd", "ystd", "scale", "shift", "seed", "mask", "p", "bits")


def adain_up_drop(x, cond, lw, lb, eps, p_drop, seed, mask, x_bcast=False, epoch=None, stats=None):
    """AdaIN(x, cond) -> bilinear x2 (align_corners) -> dropout.  x (B,h,w,C) -> (B,2h,2w,C).
    x_bcast: x has batch 1 and serves all cond.shape[0] conditions.
    epoch: optional device int32 scalar, the draw counter of wu_adain_up_drop_fwd_epoch.
    stats: the (sum, sum of squares) partials of x when the producing convolution already made them
    (conv3x3_stats); otherwise one pass over x computes them."""
    Bx, h, w, C = x.shape
    B = cond.shape[0] if x_bcast else Bx
    nc = cond.shape[1]
    dev = x.device
    if stats is not None:
        partial, nchunk = stats, stats.shape[1]
    else:
        nchunk = query("wu_adain_stats_chunks", h * w)
        partial = torch.empty((Bx, nchunk, C, 2), dtype=torch.float32, device=dev)
        call("wu_adain_stats", ptr(x), ptr(partial), Bx, h * w, C, stream())
    st = AdaINState()
    buf = torch.empty((5, B, C), dtype=torch.float32, device=dev)
    st.mean, st.rstd, st.ystd, st.scale, st.shift = buf[0], buf[1], buf[2], buf[3], buf[4]
    st.seed, st.mask, st.p = int(seed), mask, float(p_drop)
    call("wu_adain_style_fwd_n", ptr(cond), ptr(lw), ptr(lb), ptr(partial), ptr(st.mean), ptr(st.rstd),
         ptr(st.ystd), ptr(st.scale), ptr(st.shift), B, C, nc, h * w, nchunk, float(eps), int(x_bcast),
         stream())
    u = _act(B, 2 * h, 2 * w, C, x)
    st.bits = (torch.empty((B, 2 * h, 2 * w, C // 8), dtype=torch.uint8, device=dev)
               if st.p > 0 else None)
    call("wu_adain_up_drop_fwd_epoch", ptr(x), ptr(st.scale), ptr(st.shift), ptr(u), ptr(st.bits), B, h,
         w, C, st.p, st.seed, ptr(epoch), ptr(mask), int(x_bcast), stream())
    return u, st


def adain_apply(x, cond, lw, lb, eps):
    """AdaIN(x, cond) on its own (utils.py:41-51): x (B,h,w,C) NHWC bf16 -> same shape."""
    B, h, w, C = x.shape
    nc = cond.shape[1]
    dev = x.device
    nchunk = query("wu_adain_stats_chunks", h * w)
    partial = torch.empty((B, nchunk, C, 2), dtype=torch.float32, device=dev)
    call("wu_adain_stats", ptr(x), ptr(partial), B, h * w, C, stream())
    buf = torch.empty((5, B, C), dtype=torch.float32, device=dev)
    call("wu_adain_style_fwd", ptr(cond), ptr(lw), ptr(lb), ptr(partial), ptr(buf[0]), ptr(buf[1]),
         ptr(buf[2]), ptr(buf[3]), ptr(buf[4]), B, C, nc, h * w, float(eps), 0, stream())
    out = torch.empty_like(x)
    call("wu_adain_apply", ptr(x), ptr(buf[3]), ptr(buf[4]), ptr(out), B, h * w, C, stream())
    return out


def adain_up_drop_bwd(gu, x, cond, lw, lb, st, dlw=None, dlb=None):
    """-> (gx masked by relu'(x), dlw [4C][nc], dlb [4C])."""
    B, h, w, C = x.shape
    nc = cond.shape[1]
    dev = x.device
    nchunk = query("wu_adain_bwd_chunks", h, w, C)
    partial = torch.empty((B, nchunk, C, 2), dtype=torch.float32, device=dev)
    gz = torch.empty_like(x)
    call("wu_adain_up_drop_bwd", ptr(gu), ptr(x), ptr(st.mean), ptr(st.rstd), ptr(gz), ptr(partial),
         B, h, w, C, st.p, ptr(st.bits), stream())
    kk = torch.empty((5, B, C), dtype=torch.float32, device=dev)  # k1, k2, coef[3]
    gh = torch.empty((B, 4 * C), dtype=torch.float32, device=dev)
    if dlw is None:
        dlw = torch.empty((4 * C, nc), dtype=torch.float32, device=dev)
    if dlb is None:
        dlb = torch.empty((4 * C,), dtype=torch.float32, device=dev)
    call("wu_adain_style_bwd", ptr(cond), ptr(lw), ptr(lb), ptr(partial), nchunk, ptr(st.ystd),
         ptr(st.mean), ptr(st.rstd), ptr(kk[0]), ptr(kk[1]), ptr(kk[2]), ptr(gh), ptr(dlw), ptr(dlb), B, C,
         nc, h * w, stream())
    gx = torch.empty_like(x)
    call("wu_adain_bwd_apply", ptr(gz), ptr(x), ptr(kk[2]), ptr(gx), B, h * w, C, stream())
    return gx, dlw, dlb


class _BiasAct(torch.autograd.Function):
    """y = leaky_relu(x + bias, slope) in place on a channels_last bf16 (B, C, H, W) tensor (== NHWC
    memory), with a one-pass backward (masked gradient + deterministic bias gradient)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, bias, slope):
        B, C, H, W = x.shape
        call("wu_bias_act_fwd", ptr(x), ptr(bias), float(slope), B * H * W, C, stream())
        ctx.mark_dirty(x)
        ctx.save_for_backward(x)
        ctx.slope = float(slope)
        return x

    @staticmethod
    @on_tensor_device
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        B, C, H, W = y.shape
        gy = gy.contiguous(memory_format=torch.channels_last)
        if gy.dtype != BF16:
            gy = gy.to(BF16)
        g = torch.empty_like(gy)
        db = torch.empty((C,), dtype=torch.float32, device=y.device)
        nbytes = query("wu_bias_act_bwd_workspace_bytes", C)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=y.device)
        call("wu_bias_act_bwd", ptr(gy), ptr(y), ptr(g), ptr(db), ctx.slope, B * H * W, C, ptr(ws),
             nbytes, stream())
        return g, (db if ctx.needs_input_grad[1] else None), None


def bias_act_supported(x):
    """channels_last bf16 CUDA tensor with a power-of-two channel count >= 8."""
    C = x.shape[1]
    return (x.is_cuda and x.dtype == BF16 and x.dim() == 4 and C >= 8 and (C & (C - 1)) == 0
            and x.is_contiguous(memory_format=torch.channels_last))


def bias_act(x, bias, slope):
    return _BiasAct.apply(x, bias.float() if bias.dtype != torch.float32 else bias, slope)


class _DiscStem(torch.autograd.Function):
    """Discriminator stem: x (B,3,H,W) fp32 NCHW -> Conv(3,3)+bias -> Conv(3,64,stride 2)+bias ->
    LeakyReLU -> (B,64,H/2,W/2) bf16 channels_last.  Weights arrive spectrally normalised (the
    spectral_norm hook's output), so autograd continues into weight_orig / sigma on the torch side."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, w0, b0, w1, b1, slope):
        B, _, H, W = x.shape
        dev = x.device
        x = x.contiguous().float()
        w0c, w1c = w0.detach().contiguous().float(), w1.detach().contiguous().float()
        h1 = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        call("wu_conv3to3_fprop", ptr(x), ptr(w0c), ptr(b0), ptr(h1), B, H, W, stream())
        c1 = torch.empty((B, H // 2, W // 2, 64), dtype=BF16, device=dev)
        call("wu_conv3to64_s2_fprop", ptr(h1), ptr(w1c), ptr(b1), float(slope), ptr(c1), B, H, W,
             stream())
        ctx.save_for_backward(x, h1, c1, w0c, w1c)
        ctx.slope = float(slope)
        return c1.permute(0, 3, 1, 2)  # NCHW-shaped view of NHWC memory == channels_last

    @staticmethod
    @on_tensor_device
    def backward(ctx, gy):
        x, h1, c1, w0c, w1c = ctx.saved_tensors
        B, _, H, W = x.shape
        dev = x.device
        need_x, need_w0, need_b0, need_w1, need_b1 = ctx.needs_input_grad[:5]
        gy = gy.permute(0, 2, 3, 1)  # back to NHWC
        if gy.dtype != BF16 or not gy.is_contiguous():
            gy = gy.to(BF16).contiguous()
        npix = B * (H // 2) * (W // 2)
        g = torch.empty_like(gy)
        db1 = torch.empty((64,), dtype=torch.float32, device=dev)
        nb = query("wu_bias_act_bwd_workspace_bytes", 64)
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        call("wu_bias_act_bwd", ptr(gy), ptr(c1), ptr(g), ptr(db1), ctx.slope, npix, 64, ptr(ws), nb,
             stream())
        dw1 = None
        if need_w1:
            nb = query("wu_conv3to64_s2_wgrad_workspace_bytes", B, H, W)
            ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
            dw1 = torch.empty((64, 3, 3, 3), dtype=torch.float32, device=dev)
            call("wu_conv3to64_s2_wgrad", ptr(h1), ptr(g), ptr(dw1), None, B, H, W, ptr(ws), nb,
                 stream())
        gx = dw0 = db0 = None
        if need_x or need_w0 or need_b0:
            gh = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
            nb = query("wu_conv3to64_s2_dgrad_workspace_bytes", B, H, W)
            ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
            call("wu_conv3to64_s2_dgrad", ptr(g), ptr(w1c), ptr(gh), B, H, W, ptr(ws), nb, stream())
            nb = query("wu_conv3to3_bprop_workspace_bytes")
            ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
            gx = torch.empty_like(x) if need_x else None
            if need_w0 or need_b0:
                dw0 = torch.empty((3, 3, 3, 3), dtype=torch.float32, device=dev)
                db0 = torch.empty((3,), dtype=torch.float32, device=dev)
            call("wu_conv3to3_bprop", ptr(gh), ptr(x), ptr(w0c), ptr(gx), ptr(dw0), ptr(db0), B, H, W,
                 ptr(ws), nb, stream())
        return (gx, dw0 if need_w0 else None, db0 if need_b0 else None, dw1,
                db1 if need_b1 else None, None)


def disc_stem_supported(x):
    return (x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 2 == 0
            and x.shape[3] % 8 == 0 and x.dtype == torch.float32)


def disc_stem(x, w0, b0, w1, b1, slope):
    return _DiscStem.apply(x, w0, b0, w1, b1, slope)


def conv3x3_s2(src, w_packed, bias, slope, cout):
    """3x3 / stride 2 / pad 1 convolution + bias + LeakyReLU(slope) (slope 1 = none); see
    wu_conv3x3_s2_fprop.  src (B,H,W,cin) -> (B,ceil(H/2),ceil(W/2),cout)."""
    B, H, W, cin = src.shape
    dst = _act(B, (H + 1) // 2, (W + 1) // 2, cout, src)
    call("wu_conv3x3_s2_fprop", ptr(src), cin, ptr(w_packed), ptr(bias), float(slope), ptr(dst), cout,
         B, H, W, stream())
    return dst


def conv3x3_s2_dgrad(dy, w_dgrad, cin, H, W):
    """Data gradient of the stride-2 convolution: dy (B,Ho,Wo,cout) -> dx (B,H,W,cin)."""
    B, Ho, Wo, cout = dy.shape
    assert Ho == (H + 1) // 2 and Wo == (W + 1) // 2
    dx = _act(B, H, W, cin, dy)
    call("wu_conv3x3_s2_dgrad", ptr(dy), cout, ptr(w_dgrad), ptr(dx), cin, B, H, W, stream())
    return dx


def conv3x3_s2_wgrad(src, dy, want_bias=True):
    """-> (dw fp32 [cout][cin][3][3], db fp32 [cout] or None) of the stride-2 convolution."""
    B, H, W, cin = src.shape
    cout = dy.shape[3]
    nbytes = query("wu_conv3x3_s2_wgrad_workspace_bytes", cin, cout, B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dy.device)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dy.device)
    db = torch.empty((cout,), dtype=torch.float32, device=dy.device) if want_bias else None
    call("wu_conv3x3_s2_wgrad", ptr(src), cin, ptr(dy), cout, B, H, W, ptr(dw), ptr(db), ptr(ws),
         nbytes, stream())
    return dw, db


class _DiscBlock(torch.autograd.Function):
    """One discriminator block (nets.py:26-33) with cin >= 64 on the tcgen05 kernels:
    x (B,cin,H,W) bf16 channels_last -> Conv3x3(cin,cin)+bias -> Conv3x3(cin,cout,stride 2)+bias ->
    LeakyReLU -> (B,cout,H/2,W/2) bf16 channels_last.  Weights arrive spectrally normalised (fp32);
    their gradients flow back into the spectral-norm graph (weight_orig, sigma) on the torch side."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, w0, b0, w1, b1, slope, packed):
        xn = x.permute(0, 2, 3, 1)  # NHWC view of channels_last memory
        if xn.dtype != BF16 or not xn.is_contiguous():
            xn = xn.to(BF16).contiguous()
        B, H, W, cin = xn.shape
        cout = w1.shape[0]
        if packed is not None:  # bf16 operand layouts written by the fused spectral norm (w0 / w1
            (w0f, w0d), (w1f, w1d) = packed  # are autograd handles only)
        else:
            w0f, w0d = pack_conv3x3_weights(w0.detach().float().contiguous())
            w1f, w1d = pack_conv3x3_weights(w1.detach().float().contiguous())
        h = conv3x3(xn, None, w0f, b0.detach().float(), False, None, cin)  # no activation (nets.py:28-29)
        y = conv3x3_s2(h, w1f, b1.detach().float(), slope, cout)
        ctx.save_for_backward(xn, h, y, w0d, w1d)
        ctx.slope = float(slope)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @on_tensor_device
    def backward(ctx, gy):
        xn, h, y, w0d, w1d = ctx.saved_tensors
        B, H, W, cin = xn.shape
        cout = y.shape[3]
        dev = xn.device
        need_x, need_w0, need_b0, need_w1, need_b1 = ctx.needs_input_grad[:5]
        gy = gy.permute(0, 2, 3, 1)
        if gy.dtype != BF16 or not gy.is_contiguous():
            gy = gy.to(BF16).contiguous()
        g = torch.empty_like(gy)
        db1 = torch.empty((cout,), dtype=torch.float32, device=dev)
        nb = query("wu_bias_act_bwd_workspace_bytes", cout)
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        call("wu_bias_act_bwd", ptr(gy), ptr(y), ptr(g), ptr(db1), ctx.slope, B * y.shape[1] * y.shape[2],
             cout, ptr(ws), nb, stream())
        dw1 = conv3x3_s2_wgrad(h, g, want_bias=False)[0] if need_w1 else None
        gx = dw0 = db0 = None
        if need_x or need_w0 or need_b0:
            gh = conv3x3_s2_dgrad(g, w1d, cin, H, W)
            if need_w0 or need_b0:
                dw0, db0 = conv3x3_wgrad(xn, None, gh, want_bias=True)
            if need_x:
                gx = conv3x3(gh, None, w0d, None, False, None, cin).permute(0, 3, 1, 2)
        return (gx, dw0 if need_w0 else None, db0 if need_b0 else None, dw1,
                db1 if need_b1 else None, None, None)


def disc_block_supported(x):
    """bf16 channels_last CUDA activation whose channel count the tcgen05 kernels take."""
    C = x.shape[1]
    return (x.is_cuda and x.dim() == 4 and x.dtype == BF16 and C in (64, 128, 256)
            and x.shape[2] >= 2 and x.shape[3] >= 2
            and x.is_contiguous(memory_format=torch.channels_last))


def disc_block(x, w0, b0, w1, b1, slope, packed=None):
    """packed: ((w0_fprop, w0_dgrad), (w1_fprop, w1_dgrad)) when the caller already holds the bf16
    operand layouts of the two weights (then w0 / w1 only carry the autograd graph)."""
    return _DiscBlock.apply(x, w0, b0, w1, b1, slope, packed)


class _L1PerSample(torch.autograd.Function):
    """d[s] = mean |a[s] - b[s]| per sample in one pass (and one pass backward, gradient w.r.t. `a`)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, a, b):
        B = a.shape[0]
        n = a[0].numel()
        d = torch.empty((B,), dtype=torch.float32, device=a.device)
        nb = query("wu_l1_per_sample_workspace_bytes", B)
        ws = torch.empty((nb,), dtype=torch.uint8, device=a.device)
        call("wu_l1_per_sample_fwd", ptr(a), ptr(b), ptr(d), B, n, ptr(ws), nb, stream())
        ctx.save_for_backward(a, b)
        return d

    @staticmethod
    @on_tensor_device
    def backward(ctx, gd):
        a, b = ctx.saved_tensors
        ga = torch.empty_like(a)
        call("wu_l1_per_sample_bwd", ptr(a), ptr(b), ptr(gd.contiguous().float()), ptr(ga), a.shape[0],
             a[0].numel(), stream())
        return ga, None


def l1_per_sample_supported(a, b):
    return (a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
            and a.shape == b.shape and a.is_contiguous() and b.is_contiguous() and not b.requires_grad
            and a[0].numel() % 4 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)


def l1_per_sample(a, b):
    """mean over all but the batch dimension of |a - b| (gradient flows to `a` only)."""
    return _L1PerSample.apply(a, b)


def nchw_to_nhwc(x):
    """fp32 NCHW -> bf16 NHWC."""
    B, C, H, W = x.shape
    dst = _act(B, H, W, C, x)
    call("wu_nchw_f32_to_nhwc_bf16", ptr(x), ptr(dst), B, C, H, W, stream())
    return dst


def nhwc_to_nchw(x):
    """bf16 NHWC -> fp32 NCHW."""
    B, H, W, C = x.shape
    dst = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    call("wu_nhwc_bf16_to_nchw_f32", ptr(x), ptr(dst), B, C, H, W, stream())
    return dst


def launch_count():
    return int(query("wu_launch_count"))


__all__ = [n for n in dir() if not n.startswith("_")] + ["_lib"]
