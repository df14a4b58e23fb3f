"""Drop-in for the reference's cunet.py: same class name, constructor argument, sub-module names
(hence state_dict keys) and forward(x, c) signature (reference cunet.py:7-82), with the arithmetic
running in hand-written sm_100a kernels behind include/wu_b200.h.
"""
import torch
import torch.nn as nn

try:  # package import (weather_unet_b200.cunet)
    from .utils import AdaIN, HalfDropout, BatchNorm  # noqa: F401
    from .nets import r_double_conv
    from ._generator import generator_forward, PackedWeights, PACKED_NAMES
except ImportError:  # flat import with this directory on sys.path (`from cunet import ...`)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    from weather_unet_b200.utils import AdaIN, HalfDropout, BatchNorm  # noqa: F401
    from weather_unet_b200.nets import r_double_conv
    from weather_unet_b200._generator import generator_forward, PackedWeights, PACKED_NAMES


class Conditional_UNet(nn.Module):
    """Conditional U-Net generator.  Sub-modules exist to carry the parameters under the
    reference's names (39 state tensors, including the unused ``adain*.emb.weight``,
    utils.py:32); ``forward`` does not call them one by one but runs the fused kernel schedule in
    ``_generator.py``."""

    def init_weight(self, std=0.2):
        # reference cunet.py:9-16 (never called there: the call at cunet.py:41 is commented out)
        for m in self.modules():
            kind = type(m).__name__
            if "Conv" in kind:
                m.weight.data.normal_(0., std)
            elif "Linear" in kind:
                m.weight.data.normal_(1., std)
                m.bias.data.fill_(0)

    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes
        # registration order follows cunet.py:21-40 so that state_dict() iterates identically
        self.dconv_down1 = r_double_conv(3, 64)
        self.dconv_down2 = r_double_conv(64, 128)
        self.dconv_down3 = r_double_conv(128, 256)
        self.dconv_down4 = r_double_conv(256, 512)

        self.upsample = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.maxpool = nn.MaxPool2d(2)
        self.dropout = nn.Dropout(p=0.3)

        self.adain3 = AdaIN(512, num_classes=num_classes)
        self.adain2 = AdaIN(256, num_classes=num_classes)
        self.adain1 = AdaIN(128, num_classes=num_classes)

        self.dconv_up3 = r_double_conv(256 + 512, 256)
        self.dconv_up2 = r_double_conv(128 + 256, 128)
        self.dconv_up1 = r_double_conv(64 + 128, 64)

        self.conv_last = nn.Conv2d(64, 3, 1)
        self.activation = nn.Tanh()
        self._packed = PackedWeights()  # derived bf16 weights: not parameters, not persistent
        self._grad_sink = None  # data-parallel gradient buckets (train_step.GradBuckets), if any
        self._drop_seed, self._drop_epoch = 0, None  # see use_device_dropout_counter

    def forward(self, x, c, dropout_masks=None, seed=None, _keep_acts=None):
        """x: (B, 3, H, W) float in [-1, 1], H and W divisible by 8; c: (B, num_classes) float.
        Returns (B, 3, H, W) fp32 in (-1, 1).

        Extensions over the reference signature (all optional, used by the parity tests):
        ``dropout_masks`` = three uint8 NHWC keep-masks to inject instead of the Philox stream,
        ``seed`` = explicit dropout seed.  Dropout follows ``self.training`` like nn.Dropout."""
        return generator_forward(self, x, c, dropout_masks=dropout_masks, seed=seed,
                                 keep_acts=_keep_acts)

    def use_device_dropout_counter(self, enable=True):
        """Draw the dropout masks (cunet.py:61,68,75) from one base seed (taken from torch's host RNG
        now) plus a counter that lives on the device and advances with every training forward,
        instead of one host RNG draw per forward.  Statistically the same stream family; required
        when forward passes are replayed from a CUDA graph (a host-drawn seed would be frozen into
        the graph and every replay would reuse one mask).  The 16-bit counter field wraps after
        65 536 forwards, after which masks repeat with a different image/mask pairing only if the
        data repeats in lock-step."""
        if enable:
            self._drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self._drop_epoch = torch.zeros((), dtype=torch.int32, device=self.conv_last.weight.device)
        else:
            self._drop_epoch = None
        return self

    def packed_weight_names(self):
        """Names of the parameters that have derived bf16 operand copies in `self._packed` (what
        optim.FusedAdam.attach_packed maintains)."""
        return PACKED_NAMES

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._packed.invalidate()  # copy_() bumps the version counters anyway; belt and braces
        return out

    def _apply(self, fn, *args, **kwargs):
        self._packed.clear()  # .cuda()/.to() move the master weights: drop derived copies
        out = super()._apply(fn, *args, **kwargs)
        if self._drop_epoch is not None:
            self._drop_epoch = fn(self._drop_epoch)
        return out
