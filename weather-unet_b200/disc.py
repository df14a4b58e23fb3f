"""Drop-in for the reference's disc.py (disc.py:8-38): spectral-norm projection discriminator.
It is part of the measured G+D step but not of the generator hot path, so it runs
on PyTorch until the generator rows meet their bar.  Same sub-module names and state_dict keys
(`conv{1..4}.{0,1}.weight_orig/_u/_v`, `l.*`, `embed.*`), same init (disc.py:16-25), same return
list [out, c1, c2, c3, c4]."""
import numpy as np
import torch
import torch.nn as nn

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

    def forward(self, x, c=None):
        feats = []
        h = x
        for i in range(1, 5):
            h = getattr(self, f"conv{i}")(h)
            feats.append(h)
        pooled = feats[-1].sum(dim=(2, 3))  # global SUM pool (disc.py:32)
        out = self.l(pooled)
        proj = self.embed(c)  # like the reference, c=None fails here (disc.py:34)
        out = out + (proj * pooled).sum(dim=1, keepdim=True)
        return [out] + feats
