"""The cUNet generator (cunet.py:43-82 of the reference) as ONE autograd node over the C ABI.

Forward and backward are explicit kernel schedules: NHWC bf16 activations, fp32 master parameters
in the reference's state_dict layout, bf16 packed weight copies, fp32 gradients.  The skip
concatenations (cunet.py:62,69,76) are never materialised: the decoder convolutions walk two
sources in their K loop, and their data-gradient is produced as two channel slices.
"""
import torch

from . import _ops as K
from ._lib import require_device

# parameter order of the flat argument list of _CUNetFn
BLOCKS = ("dconv_down1", "dconv_down2", "dconv_down3", "dconv_down4", "dconv_up3", "dconv_up2",
          "dconv_up1")
ADAINS = ("adain3", "adain2", "adain1")


def param_names():
    names = []
    for b in BLOCKS:
        names += [f"{b}.0.weight", f"{b}.0.bias", f"{b}.2.weight", f"{b}.2.bias"]
    for a in ADAINS:
        names += [f"{a}.l1.weight", f"{a}.l1.bias"]
    names += ["conv_last.weight", "conv_last.bias"]
    return names


PARAM_NAMES = param_names()


PACKED_NAMES = tuple(f"{b}.{i}.weight" for b in BLOCKS for i in (0, 2) if (b, i) != ("dconv_down1", 0))


class PackedWeights:
    """bf16 K-major copies (w_fprop, w_dgrad) of the 13 tensor-core convolution weights.

    One persistent pair of buffers per weight (stable addresses: the optimiser's device tables and
    captured CUDA graphs point at them).  A copy is valid for one (storage, version) of its fp32
    master: any in-place change of the master bumps its version counter and the next `get` repacks,
    unless the writer has already rewritten the copies itself and said so (`mark_fresh`: the fused
    Adam launch does, optim.py)."""

    def __init__(self):
        self._buf = {}   # name -> (wf, wd)
        self._key = {}   # name -> (data_ptr, version, device) the copies were made from

    @staticmethod
    def _key_of(w):
        return (w.data_ptr(), w._version, w.device)

    def buffers(self, name, w):
        """The (w_fprop, w_dgrad) buffers of `name`, allocated for `w`'s shape / device if needed
        (contents unspecified until `get` or a writer fills them)."""
        hit = self._buf.get(name)
        cout, cin = w.shape[0], w.shape[1]
        if hit is None or hit[0].device != w.device or hit[0].shape != (cout, 9 * cin):
            hit = (torch.empty((cout, 9 * cin), dtype=torch.bfloat16, device=w.device),
                   torch.empty((cin, 9 * cout), dtype=torch.bfloat16, device=w.device))
            self._buf[name] = hit
            self._key.pop(name, None)
        return hit

    def get(self, name, w):
        wf, wd = self.buffers(name, w)
        key = self._key_of(w)
        if self._key.get(name) != key:
            with torch.no_grad():
                K.pack_conv3x3_weights_into(w.detach(), wf, wd)
            self._key[name] = key
        return wf, wd

    def mark_fresh(self, name, w):
        """The caller has just rewritten the copies of `name` from the current contents of `w`."""
        if name in self._buf:
            self._key[name] = self._key_of(w)

    def invalidate(self):
        """Forget which masters the copies were made from (buffers and their addresses stay)."""
        self._key.clear()

    def clear(self):
        self._buf.clear()
        self._key.clear()


class _CUNetFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, c, opts, *params):
        P = dict(zip(PARAM_NAMES, params))
        packed = opts["packed"]
        p_drop = opts["p"] if opts["training"] else 0.0
        masks = opts.get("masks") or (None, None, None)
        seed = opts.get("seed", 0)
        epoch = opts.get("epoch")
        eps = opts["eps"]

        def wf(name):
            return packed.get(name, P[name])[0]

        def block(src0, src1, name, stats=False):
            """r_double_conv (nets.py:18-24).  stats: the second convolution also emits AdaIN's
            sums of its output (the tensor goes to an AdaIN site next, cunet.py:59,66,73)."""
            cout = P[f"{name}.0.weight"].shape[0]
            a = K.conv3x3(src0, src1, wf(f"{name}.0.weight"), P[f"{name}.0.bias"], True, None, cout)
            if stats:
                b, sums = K.conv3x3_stats(a, None, wf(f"{name}.2.weight"), P[f"{name}.2.bias"], cout)
                return a, b, sums
            b = K.conv3x3(a, None, wf(f"{name}.2.weight"), P[f"{name}.2.bias"], True, None, cout)
            return a, b

        # encoder (cunet.py:45-54)
        a1 = K.conv_first(x, P["dconv_down1.0.weight"], P["dconv_down1.0.bias"])
        # the first two pooling layers come out of the producing convolution's epilogue (cunet.py:46,49)
        conv1, p1 = K.conv3x3_pool(a1, wf("dconv_down1.2.weight"), P["dconv_down1.2.bias"], 64)
        d2a = K.conv3x3(p1, None, wf("dconv_down2.0.weight"), P["dconv_down2.0.bias"], True, None, 128)
        conv2, p2 = K.conv3x3_pool(d2a, wf("dconv_down2.2.weight"), P["dconv_down2.2.bias"], 128)
        d3a, conv3 = block(p2, None, "dconv_down3")
        p3 = K.maxpool2(conv3)
        d4a, x4, sums4 = block(p3, None, "dconv_down4", stats=True)
        # decoder (cunet.py:59-78)
        u3, st3 = K.adain_up_drop(x4, c, P["adain3.l1.weight"], P["adain3.l1.bias"], eps[0], p_drop,
                                  seed, masks[0], epoch=epoch, stats=sums4)
        up3a, up3b, sums3 = block(u3, conv3, "dconv_up3", stats=True)
        u2, st2 = K.adain_up_drop(up3b, c, P["adain2.l1.weight"], P["adain2.l1.bias"], eps[1], p_drop,
                                  seed + 1, masks[1], epoch=epoch, stats=sums3)
        up2a, up2b, sums2 = block(u2, conv2, "dconv_up2", stats=True)
        u1, st1 = K.adain_up_drop(up2b, c, P["adain1.l1.weight"], P["adain1.l1.bias"], eps[2], p_drop,
                                  seed + 2, masks[2], epoch=epoch, stats=sums2)
        # last block: its second convolution also applies conv_last + tanh from registers (cunet.py:78-82)
        up1a = K.conv3x3(u1, conv1, wf("dconv_up1.0.weight"), P["dconv_up1.0.bias"], True, None, 64)
        up1b, y = K.conv3x3_last(up1a, wf("dconv_up1.2.weight"), P["dconv_up1.2.bias"],
                                 P["conv_last.weight"], P["conv_last.bias"])

        # x, c, the output and the parameters go through save_for_backward so that autograd notices
        # an in-place change between forward and backward (e.g. an optimiser step in between)
        ctx.save_for_backward(x, c, y, *params)
        ctx.acts = dict(a1=a1, conv1=conv1, p1=p1, d2a=d2a, conv2=conv2, p2=p2, d3a=d3a,
                        conv3=conv3, p3=p3, d4a=d4a, x4=x4, u3=u3, up3a=up3a, up3b=up3b, u2=u2,
                        up2a=up2a, up2b=up2b, u1=u1, up1a=up1a, up1b=up1b)
        ctx.sts = (st3, st2, st1)
        ctx.packed = packed
        ctx.grad_sink = opts.get("grad_sink")
        if opts.get("keep_acts") is not None:
            opts["keep_acts"].update(ctx.acts, x=x, c=c, y=y)
        return y

    @staticmethod
    def backward(ctx, gy):
        # the autograd engine's thread may have another current device: launch where the data is
        with torch.cuda.device(gy.device):
            return _CUNetFn._backward(ctx, gy)

    @staticmethod
    def _backward(ctx, gy):
        x, c, y, *params = ctx.saved_tensors
        P = dict(zip(PARAM_NAMES, params))
        A, packed = dict(ctx.acts, x=x, c=c, y=y), ctx.packed
        st3, st2, st1 = ctx.sts
        sink = ctx.grad_sink

        class _Grads(dict):
            """Parameter gradients as they are produced.  In data-parallel mode each one is written
            straight into its flat all-reduce bucket and reported ready, so the bucket's NCCL
            all-reduce starts while the rest of the backward pass is still running."""

            def __setitem__(self, name, t):
                if sink is not None and t is not None and name in sink.where:
                    v = sink.grad_view(name)
                    if t.data_ptr() != v.data_ptr():  # not produced in place (see `dst`)
                        v.copy_(t.view_as(P[name]))
                    sink.ready(name)
                    t = None
                dict.__setitem__(self, name, t)

        G = _Grads()

        def dst(name):
            """Where the kernel should write the gradient of `name`: its slot in the all-reduce
            bucket when there is one (no copy kernel afterwards), else a fresh tensor (None)."""
            if sink is not None and name in sink.where:
                v = sink.grad_view(name)
                if v.is_contiguous():
                    return v
            return None

        def wd(name):
            return packed.get(name, P[name])[1]

        def plain_block_bwd(name, src, a, g_b, mask_src):
            """Backward of r_double_conv with a single source.  g_b is masked by relu'(b).
            Returns the gradient at `src`, masked by relu'(mask_src) when given."""
            cin = P[f"{name}.0.weight"].shape[1]
            cout = P[f"{name}.0.weight"].shape[0]
            G[f"{name}.2.weight"], G[f"{name}.2.bias"] = K.conv3x3_wgrad(
                a, None, g_b, dw=dst(f"{name}.2.weight"), db=dst(f"{name}.2.bias"))
            g_a = K.conv3x3(g_b, None, wd(f"{name}.2.weight"), None, False, a, cout)
            G[f"{name}.0.weight"], G[f"{name}.0.bias"] = K.conv3x3_wgrad(
                src, None, g_a, dw=dst(f"{name}.0.weight"), db=dst(f"{name}.0.bias"))
            return K.conv3x3(g_a, None, wd(f"{name}.0.weight"), None, False, mask_src, cin)

        def up_block_bwd(name, u, skip, a, g_b):
            """Backward of a decoder r_double_conv fed by the virtual concat [u, skip]."""
            c0, c1 = u.shape[3], skip.shape[3]
            cout = a.shape[3]
            G[f"{name}.2.weight"], G[f"{name}.2.bias"] = K.conv3x3_wgrad(
                a, None, g_b, dw=dst(f"{name}.2.weight"), db=dst(f"{name}.2.bias"))
            g_a = K.conv3x3(g_b, None, wd(f"{name}.2.weight"), None, False, a, cout)
            G[f"{name}.0.weight"], G[f"{name}.0.bias"] = K.conv3x3_wgrad(
                u, skip, g_a, dw=dst(f"{name}.0.weight"), db=dst(f"{name}.0.bias"))
            w_d = wd(f"{name}.0.weight")  # [c0 + c1][9 * cout]: row slices are the two sources
            g_u = K.conv3x3(g_a, None, w_d[:c0], None, False, None, c0)
            g_skip = K.conv3x3(g_a, None, w_d[c0:], None, False, None, c1)
            return g_u, g_skip

        def adain_bwd(name, g_u, x_in, st):
            gx, dlw, dlb = K.adain_up_drop_bwd(g_u, x_in, c, P[f"{name}.l1.weight"],
                                               P[f"{name}.l1.bias"], st, dlw=dst(f"{name}.l1.weight"),
                                               dlb=dst(f"{name}.l1.bias"))
            G[f"{name}.l1.weight"], G[f"{name}.l1.bias"] = dlw, dlb
            return gx

        gy = gy.contiguous().float()
        g_up1b, G["conv_last.weight"], G["conv_last.bias"] = K.conv_last_tanh_bprop(
            gy, A["y"], A["up1b"], P["conv_last.weight"], dw=dst("conv_last.weight"),
            db=dst("conv_last.bias"))
        g_u1, g_skip1 = up_block_bwd("dconv_up1", A["u1"], A["conv1"], A["up1a"], g_up1b)
        g_up2b = adain_bwd("adain1", g_u1, A["up2b"], st1)
        g_u2, g_skip2 = up_block_bwd("dconv_up2", A["u2"], A["conv2"], A["up2a"], g_up2b)
        g_up3b = adain_bwd("adain2", g_u2, A["up3b"], st2)
        g_u3, g_skip3 = up_block_bwd("dconv_up3", A["u3"], A["conv3"], A["up3a"], g_up3b)
        g_x4 = adain_bwd("adain3", g_u3, A["x4"], st3)

        g_p3 = plain_block_bwd("dconv_down4", A["p3"], A["d4a"], g_x4, None)
        g_conv3 = K.maxpool2_bwd(A["conv3"], g_p3, g_skip3)
        g_p2 = plain_block_bwd("dconv_down3", A["p2"], A["d3a"], g_conv3, None)
        g_conv2 = K.maxpool2_bwd(A["conv2"], g_p2, g_skip2)
        g_p1 = plain_block_bwd("dconv_down2", A["p1"], A["d2a"], g_conv2, None)
        g_conv1 = K.maxpool2_bwd(A["conv1"], g_p1, g_skip1)

        G["dconv_down1.2.weight"], G["dconv_down1.2.bias"] = K.conv3x3_wgrad(
            A["a1"], None, g_conv1, dw=dst("dconv_down1.2.weight"), db=dst("dconv_down1.2.bias"))
        g_a1 = K.conv3x3(g_conv1, None, wd("dconv_down1.2.weight"), None, False, A["a1"], 64)
        G["dconv_down1.0.weight"], G["dconv_down1.0.bias"] = K.conv_first_wgrad(
            A["x"], g_a1, dw=dst("dconv_down1.0.weight"), db=dst("dconv_down1.0.bias"))

        ctx.acts = None
        grads = []
        for i, n in enumerate(PARAM_NAMES):
            g = G[n]
            grads.append(g.view_as(P[n]) if (g is not None and ctx.needs_input_grad[3 + i]) else None)
        # no gradient for the image or the condition (the reference never asks for them:
        # t_cls_train.py:242,272 differentiates w.r.t. the generator's parameters only)
        return (None, None, None) + tuple(grads)


def transfer_forward(module, x1, c, masks, seed, epoch=None):
    """One image x many conditions (inference/inf_1year_signals.py:98-107: every batch row is the
    same image, dataset.py:200-203): the encoder (cunet.py:45-54, 26.8 of 84.8 GFLOP) and the AdaIN
    statistics of the bottleneck run ONCE on the single image; the decoder runs per condition and
    reads the skip tensors through a batch-broadcast TMA coordinate.  Bit-identical to running the
    replicated batch.  Inference only (no autograd graph)."""
    P = {n: module.get_parameter(n).detach() for n in PARAM_NAMES}
    packed = module._packed
    p_drop = module.dropout.p if module.training else 0.0
    masks = masks or (None, None, None)

    def wf(name):
        return packed.get(name, P[name])[0]

    def block(src0, src1, name, bcast=False, stats=True):
        """-> (output, AdaIN sums of the output or None)"""
        cout = P[f"{name}.0.weight"].shape[0]
        a = K.conv3x3(src0, src1, wf(f"{name}.0.weight"), P[f"{name}.0.bias"], True, None, cout,
                      src1_bcast=bcast)
        if not stats:
            return K.conv3x3(a, None, wf(f"{name}.2.weight"), P[f"{name}.2.bias"], True, None, cout), None
        return K.conv3x3_stats(a, None, wf(f"{name}.2.weight"), P[f"{name}.2.bias"], cout)

    a1 = K.conv_first(x1, P["dconv_down1.0.weight"], P["dconv_down1.0.bias"])
    conv1, p1 = K.conv3x3_pool(a1, wf("dconv_down1.2.weight"), P["dconv_down1.2.bias"], 64)
    d2a = K.conv3x3(p1, None, wf("dconv_down2.0.weight"), P["dconv_down2.0.bias"], True, None, 128)
    conv2, p2 = K.conv3x3_pool(d2a, wf("dconv_down2.2.weight"), P["dconv_down2.2.bias"], 128)
    conv3, _ = block(p2, None, "dconv_down3", stats=False)
    x4, sums = block(K.maxpool2(conv3), None, "dconv_down4")
    u3, _ = K.adain_up_drop(x4, c, P["adain3.l1.weight"], P["adain3.l1.bias"], module.adain3.eps,
                            p_drop, seed, masks[0], x_bcast=True, epoch=epoch, stats=sums)
    h, sums = block(u3, conv3, "dconv_up3", bcast=True)
    u2, _ = K.adain_up_drop(h, c, P["adain2.l1.weight"], P["adain2.l1.bias"], module.adain2.eps,
                            p_drop, seed + 1, masks[1], epoch=epoch, stats=sums)
    h, sums = block(u2, conv2, "dconv_up2", bcast=True)
    u1, _ = K.adain_up_drop(h, c, P["adain1.l1.weight"], P["adain1.l1.bias"], module.adain1.eps,
                            p_drop, seed + 2, masks[2], epoch=epoch, stats=sums)
    a = K.conv3x3(u1, conv1, wf("dconv_up1.0.weight"), P["dconv_up1.0.bias"], True, None, 64,
                  src1_bcast=True)
    return K.conv3x3_last(a, wf("dconv_up1.2.weight"), P["dconv_up1.2.bias"], P["conv_last.weight"],
                          P["conv_last.bias"])[1]


def generator_forward(module, x, c, dropout_masks=None, seed=None, keep_acts=None):
    """Host-side checks (utils.py:42 batch assert, cunet.py H%8 requirement) + the autograd node."""
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"Conditional_UNet expects x of shape (B, 3, H, W), got {tuple(x.shape)}")
    if c.dim() != 2 or c.shape[1] != module.num_classes:
        raise ValueError(
            f"Conditional_UNet expects c of shape (B, {module.num_classes}), got {tuple(c.shape)}")
    # one image x many conditions: an explicit (1, 3, H, W) image with B conditions (extension), or
    # a batch that is a stride-0 expand() of one image — detected for free, no data comparison
    one_to_many = c.size(0) > 1 and (x.size(0) == 1 or (x.size(0) == c.size(0) and x.stride(0) == 0))
    if one_to_many and (torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())):
        one_to_many = False  # training needs per-sample activations: use the regular path
    if not one_to_many:
        assert x.size(0) == c.size(0)  # same failure mode as utils.py:42
    if x.shape[2] % 8 or x.shape[3] % 8:
        raise RuntimeError(
            f"Sizes of tensors must match: H and W must be divisible by 8, got {tuple(x.shape[2:])}")
    require_device(x)
    if c.device != x.device:
        raise RuntimeError("x and c must be on the same device")
    if x.requires_grad or c.requires_grad:
        raise RuntimeError("weather-unet_b200: gradients w.r.t. x or c are not provided "
                           "(the reference trains the generator's parameters only)")
    if one_to_many and keep_acts is None:
        x = x[:1]  # never materialise the replicated batch
    elif x.size(0) == 1 and c.size(0) > 1:
        x = x.expand(c.size(0), -1, -1, -1)
    x = x.detach().contiguous().float()
    c = c.detach().contiguous().float()
    training = module.training
    masks = None
    if dropout_masks is not None:
        masks = tuple(m.contiguous() for m in dropout_masks)
        B, H, W = c.shape[0], x.shape[2], x.shape[3]
        want = [(B, H // 4, W // 4, 512), (B, H // 2, W // 2, 256), (B, H, W, 128)]
        for m, s in zip(masks, want):
            if tuple(m.shape) != s or m.dtype != torch.uint8:
                raise ValueError(f"dropout mask must be uint8 NHWC {s}, got {m.dtype} {tuple(m.shape)}")
    epoch = None
    if seed is None and training and masks is None and getattr(module, "_drop_epoch", None) is not None:
        # device-side draw counter (Conditional_UNet.use_device_dropout_counter): one base seed, the
        # counter advances on the device with every forward — what a CUDA-graph replay needs
        seed, epoch = module._drop_seed, module._drop_epoch
        epoch.add_(1)
    elif seed is None:
        # host RNG draw (no device sync); three sites use seed, seed+1, seed+2
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
    params = [module.get_parameter(n) for n in PARAM_NAMES]
    if any(p.device != x.device for p in params):
        raise RuntimeError("Conditional_UNet: input and parameters are on different devices")
    # kernels, TMA descriptors and the stream all belong to the device that holds the tensors, which
    # need not be the caller's current device
    with torch.cuda.device(x.device):
        if one_to_many and keep_acts is None:
            return transfer_forward(module, x, c, masks, seed, epoch)
        opts = dict(training=training, p=module.dropout.p, masks=masks, seed=seed, epoch=epoch,
                    eps=(module.adain3.eps, module.adain2.eps, module.adain1.eps),
                    packed=module._packed, keep_acts=keep_acts,
                    grad_sink=getattr(module, "_grad_sink", None))
        return _CUNetFn.apply(x, c, opts, *params)
