"""Multi-tensor Adam on the sm_100a library: the whole parameter list of a model in one launch,
which also rewrites the packed bf16 operand copies of the generator's tensor-core convolutions.

Same update rule and hyper-parameters as the optimisers the reference builds
(t_cls_train.py:184-185: ``torch.optim.Adam(params, lr, betas=(0.0, 0.999), weight_decay=lr/20)``):
L2 weight decay added to the gradient, bias-corrected moments, eps 1e-8.  State lives in torch
tensors with torch.optim.Adam's own keys (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter), so a
``state_dict()`` is interchangeable with torch.optim.Adam's.  Parameters without a gradient are
skipped, like torch does.

The reference's ``g_opt.step()`` (t_cls_train.py:273) changes the weights the next
``inference(images, labels)`` (:302, :242) runs on.  Here the 13 tensor-core convolutions read
derived bf16 copies (``_generator.PackedWeights``), so the update has to reach them too:
``attach_packed(module)`` makes the same launch emit ``w_fprop`` / ``w_dgrad`` from the updated
masters.  Whether or not a module is attached, every updated parameter's
version counter is bumped, so any cache keyed on it notices the raw-pointer write.
"""
import struct

import torch

from ._lib import call, stream

CHUNK = 8192  # elements per CTA (ordinary tensors)
PACK_CO, PACK_CI = 16, 64  # tile of a packed 3x3 weight per CTA (wu_adam_pack_tile)


def build_tables(items):
    """Byte images of the device tables of wu_adam_multi (include/wu_b200.h).

    items: list of dicts with integer fields p, g, m, v, n and optionally wf, wd, cout, cin.
    Returns (tensor_records, chunk_records, n_chunks).  Pure host code (unit-tested on CPU)."""
    trec, crec = bytearray(), bytearray()
    n_chunks = 0
    for ti, it in enumerate(items):
        wf = it.get("wf") or 0
        trec += struct.pack("<QQQQqQQii", it["p"], it["g"], it["m"], it["v"], it["n"], wf,
                            it.get("wd") or 0, it.get("cout", 0), it.get("cin", 0))
        if wf:
            cout, cin = it["cout"], it["cin"]
            if cout % PACK_CO or cin % PACK_CI or it["n"] != cout * cin * 9:
                raise ValueError(f"packed weight must be [cout][cin][3][3] with cout % {PACK_CO} == 0 "
                                 f"and cin % {PACK_CI} == 0, got cout={cout} cin={cin} n={it['n']}")
            for tile in range((cout // PACK_CO) * (cin // PACK_CI)):
                crec += struct.pack("<iiq", ti, 0, tile)
                n_chunks += 1
        else:
            for start in range(0, it["n"], CHUNK):
                crec += struct.pack("<iiq", ti, min(CHUNK, it["n"] - start), start)
                n_chunks += 1
    return bytes(trec), bytes(crec), n_chunks


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 device_step=False):
        # torch.optim.Adam's group keys (so that state_dict()s are interchangeable both ways); the
        # options this implementation does not have are pinned to torch's defaults
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                        maximize=False, foreach=None, capturable=False, differentiable=False,
                        fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._packed = None       # _generator.PackedWeights whose copies this optimiser maintains
        self._packed_of = {}      # id(param) -> name in that cache
        self._tables = {}         # group index -> (key, tensors, chunks, n); never serialised
        # device_step: keep the step count on the device (a CUDA-graph replay then advances it);
        # the host mirror in state[p]["step"] is advanced in step() either way.
        self._device_step = bool(device_step)
        self._dev_state = {}      # group index -> int32[4] device tensor {step, bc1, rsqrt_bc2, pad}

    # ---- packed copies --------------------------------------------------------------------------
    def attach_packed(self, module):
        """Maintain `module._packed` (the bf16 operand copies of its 3x3 convolution weights) from
        inside the update.  `module` must expose `_packed` and `packed_weight_names()`."""
        self._packed = module._packed
        self._packed_of = {}
        names = set(module.packed_weight_names())
        for n, p in module.named_parameters():
            if n in names:
                self._packed_of[id(p)] = n
        self._tables.clear()
        return self

    # ---- state ----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()      # moments were replaced: the cached pointer tables are stale
        self._dev_state.clear()

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__.setdefault("_packed", None)
        self.__dict__.setdefault("_packed_of", {})
        self.__dict__.setdefault("_device_step", False)
        self._tables = {}
        self._dev_state = {}

    def _get_tables(self, gi, items, dev):
        """Device-side pointer / chunk tables; rebuilt only when a pointer changed (autograd
        allocated new .grad tensors, moments were reloaded, packed copies were reallocated)."""
        recs = []
        for p in items:
            st = self.state[p]
            it = dict(p=p.data_ptr(), g=p.grad.data_ptr(), m=st["exp_avg"].data_ptr(),
                      v=st["exp_avg_sq"].data_ptr(), n=p.numel())
            name = self._packed_of.get(id(p))
            if name is not None:
                wf, wd = self._packed.buffers(name, p)
                it.update(wf=wf.data_ptr(), wd=wd.data_ptr(), cout=p.shape[0], cin=p.shape[1])
            recs.append(it)
        key = tuple(tuple(sorted(r.items())) for r in recs)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2], hit[3]
        trec, crec, n = build_tables(recs)
        tens = torch.frombuffer(bytearray(trec), dtype=torch.uint8).pin_memory().to(dev, non_blocking=True)
        chk = torch.frombuffer(bytearray(crec), dtype=torch.uint8).pin_memory().to(dev, non_blocking=True)
        self._tables[gi] = (key, tens, chk, n)
        return tens, chk, n

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            items = [p for p in group["params"] if p.grad is not None]
            if not items:
                continue
            dev = items[0].device
            steps = set()
            for p in items:
                span = 1 + sum((n - 1) * st_ for n, st_ in zip(p.shape, p.stride()))
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32
                        and span == p.numel() and p.grad.stride() == p.stride() and p.device == dev):
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters on one device that are dense in "
                                       "memory, with gradients of the same strides, are required")
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)  # torch.optim.Adam's key / type
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                steps.add(int(st["step"]))
            b1, b2 = group["betas"]
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay"):
                raise RuntimeError("FusedAdam: amsgrad / maximize / decoupled weight decay are not implemented")
            with torch.cuda.device(dev):
                if len(steps) > 1:
                    # parameters that joined late carry their own bias correction (torch semantics):
                    # one launch per distinct step count
                    parts = [[p for p in items if int(self.state[p]["step"]) == s] for s in sorted(steps)]
                else:
                    parts = [items]
                for pi, part in enumerate(parts):
                    tens, chk, n = self._get_tables((gi, pi, len(parts)), part, dev)
                    step_no = int(self.state[part[0]]["step"])
                    dstate = None
                    if self._device_step and len(parts) == 1:
                        dstate = self._dev_state.get(gi)
                        if dstate is None:
                            dstate = torch.zeros(4, dtype=torch.int32, device=dev)
                            dstate[0] = step_no - 1
                            self._dev_state[gi] = dstate
                    call("wu_adam_multi", tens.data_ptr(), chk.data_ptr(), n, float(group["lr"]),
                         float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                         step_no, None if dstate is None else dstate.data_ptr(), stream())
            for p in items:
                # the kernel wrote through raw pointers: tell autograd / version-keyed caches
                torch.autograd.graph.increment_version(p)
                name = self._packed_of.get(id(p))
                if name is not None:
                    self._packed.mark_fresh(name, p)
        return loss
