"""Drop-in for the reference's disc.py (disc.py:8-38): spectral-norm projection discriminator
(SURVEY §8 f1).  Same sub-module names and state_dict keys (`conv{1..4}.{0,1}.weight_orig/_u/_v`,
`l.*`, `embed.*`), same init (disc.py:16-25), same return list [out, c1, c2, c3, c4].

Two paths only.  On a CUDA fp32 image under bf16 autocast (how GDTrainStep calls it) the whole
trunk runs on the sm_100a library: spectral normalisation of all ten weights in four multi-tensor launches
(_spectral.py), the 3-channel stem on the K = 27 tensor-core / FMA kernels, conv2..conv4 on the
tcgen05 implicit-GEMM kernels (stride 1 and stride 2), hand-written backward for all of them; only
the 512-wide projection head is left to PyTorch.  Anywhere else (CPU, fp32 discriminator) the
modules run as plain PyTorch, which is also what the parity tests compare against."""
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
                and torch.get_autocast_dtype("cuda") == torch.bfloat16):
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
        """Two paths, chosen by where the data lives: the sm_100a kernels (CUDA fp32 image under
        bf16 autocast — how GDTrainStep runs it), or the reference's own arithmetic through the
        PyTorch modules (disc.py:27-38: CPU, fp32 discriminator; also what the parity tests use as
        the comparator)."""
        fast = self._fast_forward(x, c) if x.is_cuda else None
        if fast is not None:
            return fast
        feats, h = [], x
        for i in range(1, 5):
            h = getattr(self, f"conv{i}")(h)
            feats.append(h)
        pooled = h.sum(dim=(2, 3), dtype=torch.float32)  # global SUM pool (disc.py:32)
        with torch.autocast(x.device.type, enabled=False):  # projection head in fp32 (tiny GEMVs)
            out = self.l(pooled)
            proj = self.embed(c)  # like the reference, c=None fails here (disc.py:34)
            out = out + (proj * pooled).sum(dim=1, keepdim=True)
        return [out] + feats
