"""Drop-in for the reference's nets.py (nets.py:4-33): same factory names, same child indices, so
state_dict keys match (`<block>.0.weight`, `<block>.2.weight`, ...).
"""
import torch.nn as nn


def upsample_box(out_channels):
    # dead code in the reference (nets.py:4-8); kept importable
    return nn.Sequential(
        nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
        nn.BatchNorm2d(out_channels, affine=False))


def double_conv(in_channels, out_channels):
    # dead code in the reference (nets.py:10-16); kept importable
    return nn.Sequential(
        nn.Conv2d(in_channels, in_channels, 3, padding=1),
        nn.Conv2d(in_channels, out_channels, 3, padding=1, stride=2),
        nn.BatchNorm2d(out_channels, affine=False),
        nn.LeakyReLU(0.2, inplace=True))


class RDoubleConv(nn.Sequential):
    """Conv3x3 -> ReLU -> Conv3x3 -> ReLU (nets.py:18-24).  Inside Conditional_UNet the block is
    only a parameter container; called on its own it runs the same sm_100a kernels through
    NCHW fp32 <-> NHWC bf16 conversions."""

    def __init__(self, in_channels, out_channels):
        super().__init__(
            nn.Conv2d(in_channels, out_channels, 3, padding=1),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, 3, padding=1),
            nn.ReLU(inplace=True))
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
    # discriminator block (nets.py:26-33); runs on PyTorch — SURVEY §8 f1, not on the hot path yet
    return nn.Sequential(
        nn.utils.spectral_norm(nn.Conv2d(in_channels, in_channels, 3, padding=1)),
        nn.utils.spectral_norm(nn.Conv2d(in_channels, out_channels, 3, padding=1, stride=2)),
        nn.LeakyReLU(0.2, inplace=True))
