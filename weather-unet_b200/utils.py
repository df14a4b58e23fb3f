"""Drop-in for the reference's utils.py: AdaIN (utils.py:26-51) is the hot-path member; the other
names are dead code in the reference and are kept importable as plain PyTorch modules.
"""
import torch
import torch.nn as nn


class AdaIN(nn.Module):
    """Adaptive instance norm conditioned on a weather vector (utils.py:26-51).

    Parameters live under the reference's names: ``l1`` (Linear nc -> 4C) and the unused ``emb``
    (utils.py:32) which never receives a gradient but is part of the state_dict contract."""

    def __init__(self, in_channel, num_classes, eps=1e-5):
        super().__init__()
        self.num_classes = num_classes
        self.in_channel = in_channel
        self.eps = eps
        self.l1 = nn.Linear(num_classes, in_channel * 4, bias=True)
        self.emb = nn.Embedding(num_classes, num_classes)

    def forward(self, x, y):
        """Standalone use: x (B, C, H, W) fp32 NCHW, y (B, num_classes).  Inside Conditional_UNet
        the module is a parameter container and the fused AdaIN+upsample+dropout kernels run."""
        try:
            from ._blocks import adain_forward
        except ImportError:
            from weather_unet_b200._blocks import adain_forward
        return adain_forward(self, x, y)


class ConditionalNorm(nn.Module):
    # unused by the generator (utils.py:7-23)
    def __init__(self, in_channel, num_classes=5):
        super().__init__()
        self.num_classes = num_classes
        self.bn = nn.BatchNorm2d(in_channel, affine=False)
        self.embed = nn.Embedding(num_classes, in_channel * 2)
        self.embed.weight.data[:, :in_channel] = 1
        self.embed.weight.data[:, in_channel:] = 0

    def forward(self, input, class_id):
        gamma, beta = self.embed(class_id).chunk(2, 1)
        return gamma[:, :, None, None] * self.bn(input) + beta[:, :, None, None]


class BatchNorm(nn.Module):
    # unused by the generator (utils.py:54-71): per-sample standardisation over C*H*W
    def forward(self, x):
        flat = x.reshape(x.size(0), -1)
        std = (flat.var(dim=-1) + 1e-5).sqrt().view(-1, 1, 1, 1)
        mean = flat.mean(dim=-1).view(-1, 1, 1, 1)
        return (x - mean) / std


class MakeOneHot(nn.Module):
    # utils.py:74-81
    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes

    def forward(self, x):
        return nn.functional.one_hot(torch.argmax(x), self.num_classes)


class HalfDropout(nn.Module):
    # utils.py:84-95: dropout on the first half of the channels only
    def __init__(self, p=0.3):
        super().__init__()
        self.dropout = nn.Dropout(p=p)

    def forward(self, x):
        half = x.size(1) // 2
        return torch.cat([self.dropout(x[:, :half]), x[:, half:]], dim=1)


class Denormalize(object):
    # utils.py:98-109: inverse of torchvision Normalize followed by a clamp to [0, 1]
    def __init__(self, mean, std, inplace=False):
        self.mean, self.std, self.inplace = mean, std, inplace
        self.demean = [-m / s for m, s in zip(mean, std)]
        self.destd = [1 / s for s in std]

    def __call__(self, tensor):
        mean = torch.as_tensor(self.demean, dtype=tensor.dtype, device=tensor.device).view(-1, 1, 1)
        std = torch.as_tensor(self.destd, dtype=tensor.dtype, device=tensor.device).view(-1, 1, 1)
        out = tensor.sub_(mean).div_(std) if self.inplace else (tensor - mean) / std
        return torch.clamp(out, 0.0, 1.0)
