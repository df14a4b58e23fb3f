"""ctypes binding of libwu_b200.so (the C ABI declared in include/wu_b200.h).

PyTorch is plumbing here: it owns every buffer and the stream; the arithmetic is in the library.
There is no CPU / eager fallback: if the shared library is missing or the device is not sm_100,
the first call raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint64, c_ulonglong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# WU_B200_LIB: developer hook for A/B measurements of experimental builds (tools/pipe_stats.sh); the
# product always loads the in-tree library
LIB_PATH = os.environ.get("WU_B200_LIB") or os.path.join(_HERE, "lib", "libwu_b200.so")

P, I, F, U64, SZ = c_void_p, c_int, c_float, c_uint64, c_size_t

# name -> (restype, argtypes); mirrors include/wu_b200.h line by line
SIGNATURES = {
    "wu_last_error": (c_char_p, []),
    "wu_version": (I, []),
    "wu_device_check": (I, []),
    "wu_launch_count": (c_ulonglong, []),
    "wu_pack_conv3x3_weights": (I, [P, I, I, P, P, P]),
    "wu_conv3x3_fprop": (I, [P, I, P, I, P, P, I, P, P, I, I, I, I, P]),
    "wu_conv3x3_fprop_bcast": (I, [P, I, P, I, I, P, P, I, P, P, I, I, I, I, P]),
    "wu_conv3x3_stats_chunks": (I, [I, I, I]),
    "wu_conv3x3_fprop_stats": (I, [P, I, P, I, I, P, P, P, P, I, I, I, I, P]),
    "wu_conv3x3_fprop_last": (I, [P, I, P, P, P, P, P, P, I, I, I, P]),
    "wu_conv3x3_fprop_pool": (I, [P, I, P, P, P, P, I, I, I, I, P]),
    "wu_conv3x3_wgrad_workspace_bytes": (SZ, [I, I, I, I, I]),
    "wu_conv3x3_wgrad": (I, [P, I, P, I, P, I, I, I, I, P, P, P, SZ, P]),
    "wu_conv_first_fprop": (I, [P, P, P, P, I, I, I, P]),
    "wu_conv_first_wgrad_workspace_bytes": (SZ, [I, I, I]),
    "wu_conv_first_wgrad": (I, [P, P, P, P, I, I, I, P, SZ, P]),
    "wu_conv_last_tanh_fprop": (I, [P, P, P, P, I, I, I, P]),
    "wu_conv_last_tanh_bprop_workspace_bytes": (SZ, [I, I, I]),
    "wu_conv_last_tanh_bprop": (I, [P, P, P, P, P, P, P, I, I, I, P, SZ, P]),
    "wu_maxpool2_fwd": (I, [P, P, I, I, I, I, P]),
    "wu_maxpool2_bwd": (I, [P, P, P, P, I, I, I, I, P]),
    "wu_adain_stats_chunks": (I, [I]),
    "wu_adain_stats": (I, [P, P, I, I, I, P]),
    "wu_adain_style_fwd": (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, F, I, P]),
    "wu_adain_up_drop_fwd_epoch": (I, [P, P, P, P, P, I, I, I, I, F, U64, P, P, I, P]),
    "wu_adain_style_fwd_n": (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P]),
    "wu_adain_apply": (I, [P, P, P, P, I, I, I, P]),
    "wu_adain_up_drop_fwd": (I, [P, P, P, P, P, I, I, I, I, F, U64, P, I, P]),
    "wu_adain_bwd_chunks": (I, [I, I, I]),
    "wu_adain_up_drop_bwd": (I, [P, P, P, P, P, P, I, I, I, I, F, P, P]),
    "wu_adain_style_bwd": (I, [P, P, P, P, I, P, P, P, P, P, P, P, P, P, I, I, I, I, P]),
    "wu_adain_bwd_apply": (I, [P, P, P, P, I, I, I, P]),
    "wu_bias_act_fwd": (I, [P, P, F, c_longlong, I, P]),
    "wu_bias_act_bwd_workspace_bytes": (SZ, [I]),
    "wu_bias_act_bwd": (I, [P, P, P, P, F, c_longlong, I, P, SZ, P]),
    "wu_conv3to3_fprop": (I, [P, P, P, P, I, I, I, P]),
    "wu_conv3to64_s2_fprop": (I, [P, P, P, F, P, I, I, I, P]),
    "wu_conv3to64_s2_wgrad_workspace_bytes": (SZ, [I, I, I]),
    "wu_conv3to64_s2_wgrad": (I, [P, P, P, P, I, I, I, P, SZ, P]),
    "wu_conv3to64_s2_dgrad_workspace_bytes": (SZ, [I, I, I]),
    "wu_conv3to64_s2_dgrad": (I, [P, P, P, I, I, I, P, SZ, P]),
    "wu_conv3to3_bprop_workspace_bytes": (SZ, []),
    "wu_conv3to3_bprop": (I, [P, P, P, P, P, P, I, I, I, P, SZ, P]),
    "wu_conv3x3_s2_fprop": (I, [P, I, P, P, F, P, I, I, I, I, P]),
    "wu_conv3x3_s2_dgrad": (I, [P, I, P, P, I, I, I, I, P]),
    "wu_conv3x3_s2_wgrad_workspace_bytes": (SZ, [I, I, I, I, I]),
    "wu_conv3x3_s2_wgrad": (I, [P, I, P, I, I, I, I, P, P, P, SZ, P]),
    "wu_sn_parts": (I, []),
    "wu_sn_wtu_cols": (I, []),
    "wu_sn_wv_rows": (I, []),
    "wu_sn_forward": (I, [P, I, P, I, P, I, P, I, I, F, P]),
    "wu_sn_backward": (I, [P, P, I, P, I, P]),
    "wu_adam_multi": (I, [P, P, I, F, F, F, F, F, I, P, P]),
    "wu_adam_pack_tile": (I, [P, P]),
    "wu_l1_per_sample_workspace_bytes": (SZ, [I]),
    "wu_l1_per_sample_fwd": (I, [P, P, P, I, c_longlong, P, SZ, P]),
    "wu_l1_per_sample_bwd": (I, [P, P, P, P, I, c_longlong, P]),
    "wu_nchw_f32_to_nhwc_bf16": (I, [P, P, I, I, I, I, P]),
    "wu_nhwc_bf16_to_nchw_f32": (I, [P, P, I, I, I, I, P]),
}

_lib = None


class WuError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WuError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (no CPU fallback exists for the cUNet hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(SIGNATURES)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Call a status-returning entry point; raise WuError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.wu_last_error()
        raise WuError(f"{name} failed (code {rc}): {msg.decode() if msg else '?'}")


def query(name, *args):
    """Call a value-returning entry point (workspace sizes, counters)."""
    return getattr(load(), name)(*args)


def on_tensor_device(fn):
    """Decorator for autograd.Function.forward / backward: run under the CUDA device of the first
    CUDA tensor argument.  The caller's (or the autograd engine thread's) current device need not be
    the one that holds the data, and kernels, TMA descriptors and the stream all follow the current
    device."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        t = next((a for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if t is None:
            return fn(ctx, *args)
        with torch.cuda.device(t.device):
            return fn(ctx, *args)
    return wrapper


_checked_devices = set()


def require_device(t):
    """The hot path has no CPU path: tensors must live on an sm_100 CUDA device."""
    if not t.is_cuda:
        raise WuError("weather-unet_b200: input is not a CUDA tensor and there is no CPU fallback")
    idx = t.device.index
    if idx not in _checked_devices:
        with torch.cuda.device(idx):
            call("wu_device_check")
        _checked_devices.add(idx)
