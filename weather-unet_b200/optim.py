"""Multi-tensor Adam on the sm_100a library: the whole parameter list of a model in one launch.

Same update rule and hyper-parameters as the optimisers the reference builds
(t_cls_train.py:184-185: ``torch.optim.Adam(params, lr, betas=(0.0, 0.999), weight_decay=lr/20)``):
L2 weight decay added to the gradient, bias-corrected moments, eps 1e-8.  State lives in torch
tensors (exp_avg, exp_avg_sq), so ``state_dict()`` / ``load_state_dict()`` of the torch class it
derives from keep working.  Parameters without a gradient are skipped, like torch does.
"""
import struct

import torch

from ._lib import call, stream

CHUNK = 8192  # elements per CTA


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._key = None
        self._tensors = self._chunks = None
        self._n_chunks = 0

    def _tables(self, group, items, dev):
        """Device-side pointer / chunk tables; rebuilt only when a pointer changed (e.g. autograd
        allocated new .grad tensors)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in items)
        cache = group.setdefault("_wu_cache", {})
        if cache.get("key") == key:
            return cache["tensors"], cache["chunks"], cache["n"]
        trec, crec = bytearray(), bytearray()
        n_chunks = 0
        for ti, p in enumerate(items):
            st = self.state[p]
            trec += struct.pack("<QQQQq", p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                st["exp_avg_sq"].data_ptr(), p.numel())
            for start in range(0, p.numel(), CHUNK):
                crec += struct.pack("<iiq", ti, min(CHUNK, p.numel() - start), start)
                n_chunks += 1
        tens = torch.frombuffer(trec, dtype=torch.uint8).pin_memory().to(dev, non_blocking=True)
        chk = torch.frombuffer(crec, dtype=torch.uint8).pin_memory().to(dev, non_blocking=True)
        cache.update(key=key, tensors=tens, chunks=chk, n=n_chunks)
        return tens, chk, n_chunks

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            items = [p for p in group["params"] if p.grad is not None]
            if not items:
                continue
            dev = items[0].device
            for p in items:
                span = 1 + sum((n - 1) * st_ for n, st_ in zip(p.shape, p.stride()))
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32
                        and span == p.numel() and p.grad.stride() == p.stride()):
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters that are dense in memory, with "
                                       "gradients of the same strides, are required")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            group["_wu_step"] = group.get("_wu_step", 0) + 1
            tens, chk, n = self._tables(group, items, dev)
            b1, b2 = group["betas"]
            call("wu_adam_multi", tens.data_ptr(), chk.data_ptr(), n, float(group["lr"]), float(b1),
                 float(b2), float(group["eps"]), float(group["weight_decay"]), int(group["_wu_step"]),
                 stream())
        return loss
