"""Drop-in for the reference's disc.py (disc.py:8-38): spectral-norm projection discriminator
(SURVEY §8 f1).  Same sub-module names and state_dict keys (`conv{1..4}.{0,1}.weight_orig/_u/_v`,
`l.*`, `embed.*`), same init (disc.py:16-25), same return list [out, c1, c2, c3, c4].

On a CUDA fp32 image under bf16 autocast (how GDTrainStep calls it) the whole trunk runs on the
sm_100a library: spectral normalisation of all ten weights in four multi-tensor launches
(_spectral.py), the 3-channel stem on the K = 27 tensor-core / FMA kernels, conv2..conv4 on the
tcgen05 implicit-GEMM kernels (stride 1 and stride 2), hand-written backward for all of them; only
the 512-wide projection head is left to PyTorch.  Anywhere else (CPU, fp32 discriminator) the
modules run as plain PyTorch, which is also what the parity tests compare against."""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

try:
    from .nets import sn_double_conv
except ImportError:
    from weather_unet_b200.nets import sn_double_conv


class SNDisc(nn.Module):

    def __init__(self, num_classes):
        super().__init__()
        widths = (3, 64, 128, 256, 512)
        for i in range(4):
            blk = sn_double_conv(widths[i], widths[i + 1])
            setattr(self, f"conv{i + 1}", blk)
        for i in range(1, 5):
            for j in range(2):
                nn.init.xavier_uniform_(getattr(self, f"conv{i}")[j].weight, np.sqrt(2))
        self.l = nn.utils.spectral_norm(nn.Linear(512, 1))
        nn.init.xavier_uniform_(self.l.weight)
        self.embed = nn.utils.spectral_norm(nn.Linear(num_classes, 512, bias=True))
        nn.init.xavier_uniform_(self.embed.weight)

    @staticmethod
    def _sn_conv(conv, h, slope):
        """Spectral-norm convolution + bias (+ LeakyReLU when slope != 1).  On a bf16
        channels_last CUDA activation the bias add and the activation run in one sm_100a kernel
        (wu_bias_act_*) instead of separate ATen passes; anywhere else this is the plain module."""
        for hook in conv._forward_pre_hooks.values():  # spectral_norm: power iteration, W / sigma
            hook(conv, (h,))
        out = F.conv2d(h, conv.weight, None, conv.stride, conv.padding)
        try:
            from . import _ops as K
        except ImportError:
            from weather_unet_b200 import _ops as K
        if K.bias_act_supported(out):
            return K.bias_act(out, conv.bias, slope)
        out = out + conv.bias.to(out.dtype).view(1, -1, 1, 1)
        return out if slope == 1.0 else F.leaky_relu(out, slope)

    def _stem(self, x):
        """conv1 block on the sm_100a stem kernels (3-channel FMA/HBM-bound work) when the input is
        an fp32 CUDA image and bf16 autocast is on; None otherwise."""
        try:
            from . import _ops as K
        except ImportError:
            from weather_unet_b200 import _ops as K
        if not (K.disc_stem_supported(x) and torch.is_autocast_enabled()
                and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            return None
        c0, c1, act = self.conv1[0], self.conv1[1], self.conv1[2]
        with torch.autocast("cuda", enabled=False):
            for conv in (c0, c1):  # spectral norm of both weights, in fp32 like the reference
                for hook in conv._forward_pre_hooks.values():
                    hook(conv, (x,))
            return K.disc_stem(x, c0.weight.float(), c0.bias.float(), c1.weight.float(),
                               c1.bias.float(), act.negative_slope)

    @staticmethod
    def _block(blk, h):
        """conv2..conv4 blocks on the tcgen05 kernels (stride-1 implicit GEMM + the stride-2 /
        LeakyReLU kernel of wu_conv_s2.cu, hand-written backward) when `h` is a bf16 channels_last
        CUDA activation under bf16 autocast; None otherwise."""
        try:
            from . import _ops as K
        except ImportError:
            from weather_unet_b200 import _ops as K
        if not (K.disc_block_supported(h) and torch.is_autocast_enabled()
                and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            return None
        if os.environ.get("WU_DISC_TRUNK", "") == "cudnn":  # A/B measurements only
            return None
        c0, c1, act = blk[0], blk[1], blk[2]
        with torch.autocast("cuda", enabled=False):
            for conv in (c0, c1):  # spectral norm (power iteration, W / sigma) in fp32 like the reference
                for hook in conv._forward_pre_hooks.values():
                    hook(conv, (h,))
            return K.disc_block(h, c0.weight.float(), c0.bias.float(), c1.weight.float(),
                                c1.bias.float(), act.negative_slope)

    def _sn_modules(self):
        mods = []
        for i in range(1, 5):
            blk = getattr(self, f"conv{i}")
            mods += [blk[0], blk[1]]
        return mods + [self.l, self.embed]

    def _fast_forward(self, x, c):
        """The all-kernel path; None when its preconditions do not hold."""
        try:
            from . import _ops as K
            from ._spectral import FusedSpectralNorm
        except ImportError:
            from weather_unet_b200 import _ops as K
            from weather_unet_b200._spectral import FusedSpectralNorm
        if not (K.disc_stem_supported(x) and torch.is_autocast_enabled()
                and torch.get_autocast_dtype("cuda") == torch.bfloat16
                and os.environ.get("WU_DISC_TRUNK", "") != "cudnn"):
            return None
        sn = self.__dict__.get("_fused_sn")
        if sn is None:
            sn = FusedSpectralNorm(self._sn_modules())
            self.__dict__["_fused_sn"] = sn  # not a sub-module / not in the state_dict
        if not sn.supported():
            return None
        with torch.autocast("cuda", enabled=False):
            ws, packed = sn(self.training)  # W / sigma of all ten weights, fp32 like the reference
            slope = self.conv1[2].negative_slope
            h = K.disc_stem(x, ws[0], self.conv1[0].bias, ws[1], self.conv1[1].bias, slope)
            feats = [h]
            for i in range(2, 5):
                blk = getattr(self, f"conv{i}")
                if not K.disc_block_supported(h):
                    return None
                h = K.disc_block(h, ws[2 * i - 2], blk[0].bias, ws[2 * i - 1], blk[1].bias,
                                 blk[2].negative_slope, packed=(packed[2 * i - 2], packed[2 * i - 1]))
                feats.append(h)
            pooled = h.sum(dim=(2, 3), dtype=torch.float32)  # global SUM pool (disc.py:32)
            out = F.linear(pooled, ws[8], self.l.bias)
            proj = F.linear(c, ws[9], self.embed.bias)  # like the reference, c=None fails here (disc.py:34)
            out = out + (proj * pooled).sum(dim=1, keepdim=True)
        return [out] + feats

    def forward(self, x, c=None):
        fast = self._fast_forward(x, c) if x.is_cuda else None
        if fast is not None:
            return fast
        feats = []
        h = self._stem(x)
        first = 1
        if h is not None:
            feats.append(h)
            first = 2
        else:
            h = x
        for i in range(first, 5):
            blk = getattr(self, f"conv{i}")
            fast = self._block(blk, h)
            if fast is not None:
                h = fast
            else:
                h = self._sn_conv(blk[0], h, 1.0)                       # no activation in between
                h = self._sn_conv(blk[1], h, blk[2].negative_slope)     # (nets.py:26-33)
            feats.append(h)
        pooled = feats[-1].sum(dim=(2, 3), dtype=torch.float32)  # global SUM pool (disc.py:32)
        with torch.autocast(x.device.type, enabled=False):  # projection head in fp32 (tiny GEMVs)
            out = self.l(pooled)
            proj = self.embed(c)  # like the reference, c=None fails here (disc.py:34)
            out = out + (proj * pooled).sum(dim=1, keepdim=True)
        return [out] + feats
