"""Drop-in for the reference's nets.py (nets.py:4-33): same factory names, same child indices, so
state_dict keys match (`<block>.0.weight`, `<block>.2.weight`, ...).
"""
import torch.nn as nn


def _conv3x3(cin, cout, stride=1, sn=False):
    """The one convolution shape the reference uses: 3x3, padding 1, optional stride / spectral norm."""
    conv = nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1)
    return nn.utils.spectral_norm(conv) if sn else conv


def upsample_box(out_channels):
    """nets.py:4-8 (never called by the reference; kept so `from nets import *` finds it)."""
    layers = [nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
              nn.BatchNorm2d(out_channels, affine=False)]
    return nn.Sequential(*layers)


def double_conv(in_channels, out_channels):
    """nets.py:10-16 (never called by the reference; kept importable)."""
    layers = [_conv3x3(in_channels, in_channels), _conv3x3(in_channels, out_channels, stride=2),
              nn.BatchNorm2d(out_channels, affine=False), nn.LeakyReLU(0.2, inplace=True)]
    return nn.Sequential(*layers)


class RDoubleConv(nn.Sequential):
    """Conv3x3 -> ReLU -> Conv3x3 -> ReLU (nets.py:18-24).  Inside Conditional_UNet the block is
    only a parameter container; called on its own it runs the same sm_100a kernels through
    NCHW fp32 <-> NHWC bf16 conversions."""

    def __init__(self, in_channels, out_channels):
        # children 0 and 2 are the convolutions: `<block>.0.weight`, `<block>.2.weight` as in the reference
        super().__init__(_conv3x3(in_channels, out_channels), nn.ReLU(inplace=True),
                         _conv3x3(out_channels, out_channels), nn.ReLU(inplace=True))
        self.in_channels, self.out_channels = in_channels, out_channels

    def forward(self, x):
        try:
            from ._blocks import double_conv_forward
        except ImportError:
            from weather_unet_b200._blocks import double_conv_forward
        return double_conv_forward(self, x)


def r_double_conv(in_channels, out_channels):
    return RDoubleConv(in_channels, out_channels)


def sn_double_conv(in_channels, out_channels):
    """Discriminator block (nets.py:26-33): SN-conv, SN-conv stride 2, LeakyReLU(0.2).  disc.SNDisc runs
    these parameters through the sm_100a kernels (fused spectral norm, stride-2 tcgen05 convolutions);
    called as a plain module it is PyTorch's own hook-based spectral norm."""
    layers = [_conv3x3(in_channels, in_channels, sn=True),
              _conv3x3(in_channels, out_channels, stride=2, sn=True), nn.LeakyReLU(0.2, inplace=True)]
    return nn.Sequential(*layers)
