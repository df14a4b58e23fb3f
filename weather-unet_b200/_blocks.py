"""Standalone forward of the generator's building blocks (r_double_conv, AdaIN) on the sm_100a
kernels, for callers that use them outside Conditional_UNet.  Inference only: training goes
through Conditional_UNet, whose backward schedule lives in _generator.py."""
import torch

from . import _ops as K
from ._generator import PackedWeights
from ._lib import require_device


def _no_grad_only(what, *tensors):
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise RuntimeError(f"{what}: standalone use is forward-only; wrap the call in torch.no_grad() "
                           "or train through Conditional_UNet")


def double_conv_forward(module, x):
    """x fp32 NCHW (B, Cin, H, W) -> fp32 NCHW (B, Cout, H, W); Cin == 3 or Cin % 64 == 0."""
    require_device(x)
    c0, c2 = module[0], module[2]
    _no_grad_only("r_double_conv", x, c0.weight)
    if not hasattr(module, "_packed"):
        module._packed = PackedWeights()
    x = x.detach().contiguous().float()
    cin, cout = c0.weight.shape[1], c0.weight.shape[0]
    if cin == 3 and cout == 64:
        a = K.conv_first(x, c0.weight.detach(), c0.bias.detach())
    elif cin % 64 == 0:
        a = K.conv3x3(K.nchw_to_nhwc(x), None, module._packed.get("0", c0.weight)[0],
                      c0.bias.detach(), True, None, cout)
    else:
        raise ValueError(f"r_double_conv on sm_100a needs in_channels == 3 or a multiple of 64, got {cin}")
    b = K.conv3x3(a, None, module._packed.get("2", c2.weight)[0], c2.bias.detach(), True, None, cout)
    return K.nhwc_to_nchw(b)


def adain_forward(module, x, y):
    """AdaIN on its own (utils.py:41-51), without the fused upsample/dropout."""
    require_device(x)
    assert x.size(0) == y.size(0)
    _no_grad_only("AdaIN", x, y, module.l1.weight)
    xh = K.nchw_to_nhwc(x.detach().contiguous().float())
    out = K.adain_apply(xh, y.detach().contiguous().float(), module.l1.weight.detach(),
                        module.l1.bias.detach(), module.eps)
    return K.nhwc_to_nchw(out)
