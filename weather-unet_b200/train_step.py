"""One G+D training iteration of the reference's class-/estimator-conditioned trainers
(t_cls_train.py:288-312 discriminator update, :226-286 generator update, :184-185 optimisers;
t_est_train.py:214-283 is the same skeleton with real-valued conditions), restated for synthetic
or user-supplied batches, single GPU or data parallel (one process per GPU, NCCL all-reduce of the
gradient buckets overlapped with backward).

The generator and the discriminator run on the sm_100a kernels (the discriminator under bf16
autocast); the optional frozen estimator is an ordinary PyTorch module.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .ops import dis_hinge, gen_hinge, l1_loss


def _ops():
    from . import _ops as K  # deferred: the ctypes library is only needed on a GPU
    return K


class GradBuckets:
    """Flat fp32 gradient buckets with `param.grad` as views into them (no copies), all-reduced
    (sum / world) bucket by bucket on a side stream as soon as every gradient of a bucket has been
    produced.  Parameters whose gradient never arrives (adain*.emb.weight, utils.py:32) are left
    out by passing `skip`."""

    def __init__(self, named_params, group=None, bucket_bytes=8 << 20, skip=(), collective=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        # collective: run the all-reduce even in a group of one (tests of the launch / stream logic)
        self.collective = (self.world > 1) if collective is None else bool(collective)
        params = [(n, p) for n, p in named_params if p.requires_grad and n not in skip]
        params.reverse()  # gradients are produced in reverse forward order
        self.buckets, cur, cur_bytes = [], [], 0
        for n, p in params:
            cur.append((n, p))
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat, self.where, self.views, self.got = [], {}, {}, set()
        for bi, bucket in enumerate(self.buckets):
            total = sum(p.numel() for _, p in bucket)
            dev = bucket[0][1].device
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            off = 0
            for n, p in bucket:
                # same strides as the parameter (contiguous or channels_last): autograd's layout contract
                self.views[n] = torch.as_strided(flat, p.shape, p.stride(), storage_offset=off)
                p.grad = self.views[n]
                self.where[n] = bi
                off += p.numel()
            self.flat.append(flat)
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        dev = self.flat[0].device if self.flat else None
        self.comm_stream = torch.cuda.Stream(device=dev) if (dev is not None and dev.type == "cuda") else None
        self._hooks = []

    def attach_autograd_hooks(self):
        """For ordinary PyTorch modules: fire from AccumulateGrad post hooks."""
        for bucket in self.buckets:
            for n, p in bucket:
                self._hooks.append(p.register_post_accumulate_grad_hook(
                    lambda _p, n=n: self.ready(n)))

    def zero(self):
        for f in self.flat:
            f.zero_()
        for bucket in self.buckets:
            for n, p in bucket:
                p.grad = self.views[n]
        self.got = set()
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def grad_view(self, name):
        return self.views[name]

    def ready(self, name):
        """Mark one gradient as final; launch the bucket's all-reduce when it is complete."""
        bi = self.where.get(name)
        if bi is None:
            return
        self.got.add(name)
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        if self.launched[bi] or not self.collective:
            return
        self.launched[bi] = True
        flat = self.flat[bi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                # NCCL averages inside the collective: no separate scaling kernel per bucket
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # CPU / gloo (tests)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def finish(self):
        """Reduce buckets that never completed (a parameter without gradient this pass — the same
        on every rank, so the collective order still matches), then make the compute stream wait
        for the outstanding reductions.  Call before optimizer.step()."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        if self.comm_stream is not None and self.collective:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        for bucket in self.buckets:  # no gradient this pass -> grad None, as without data parallelism
            for n, p in bucket:      # (the optimiser then skips it: no weight-decay-only update)
                if n not in self.got:
                    p.grad = None


class GDTrainStep:
    """D update then G update, exactly the reference's order and losses.

    estimator: optional frozen module (B,3,H,W)->(B,nc) (t_cls_train.py:172-177).  When given, the
    weather term g_loss_w = MSE(estimator(fake), target) is added (t_cls_train.py:256,
    ops.py:37-39); when None (no checkpoint exists offline) g_loss = g_loss_adv + loss_con.
    """

    def __init__(self, G, D, lr=1e-4, estimator=None, d_autocast=True, eps_con=1e-2, group=None,
                 overlap=True, distributed=None, share_fake=False, fused_adam=None, static_grads=False,
                 bucket_bytes=8 << 20):
        self.G, self.D, self.estimator = G, D, estimator
        # share_fake=True is NOT the reference's schedule: the reference runs the generator twice per
        # iteration (t_cls_train.py:302 and :242) with two independent dropout draws; sharing one
        # forward between the D and the G update is statistically equivalent but not bitwise
        #  and saves one generator forward.  Default: faithful.
        self.share_fake = share_fake
        self.d_autocast = d_autocast
        # the sm_100a discriminator path (bf16 autocast) takes the reference's weight layout; only a
        # cuDNN-run discriminator (fp32) profits from channels_last weights
        self.d_channels_last = next(D.parameters()).is_cuda and not d_autocast
        if self.d_channels_last:
            D.to(memory_format=torch.channels_last)
        self.eps_con = eps_con  # 1e-2 supervised, 1e-7 otherwise (t_cls_train.py:259-266)
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.distributed = bool(distributed)
        self.g_buckets = self.d_buckets = None
        self._use_sink = False
        # static_grads: keep every .grad at a fixed address (views into flat buckets, written in place
        # by the generator's backward) even on one GPU — what capturing the iteration in a CUDA
        # graph needs (GraphedGDStep); it is the data-parallel bookkeeping without the collective
        if self.distributed or static_grads:
            skip = tuple(n for n, _ in G.named_parameters() if n.endswith("emb.weight"))
            coll = self.distributed and dist.is_available() and dist.is_initialized()
            self.g_buckets = GradBuckets(G.named_parameters(), group, bucket_bytes=bucket_bytes, skip=skip,
                                         collective=coll)
            self.d_buckets = GradBuckets(D.named_parameters(), group, bucket_bytes=bucket_bytes,
                                         collective=coll)
            self.d_buckets.attach_autograd_hooks()
            # the generator's backward writes its gradients straight into the buckets; the sink is
            # handed to the module only around the G update's forward (step()), so a backward
            # outside the trainer never touches the buckets
            self._use_sink = bool(overlap and hasattr(G, "_grad_sink"))
            if not self._use_sink:
                self.g_buckets.attach_autograd_hooks()
            for m in ((G, D) if self.distributed else ()):  # identical replicas to start from
                with torch.no_grad():
                    for t in list(m.parameters()) + list(m.buffers()):
                        # detach() shares the version counter (unlike .data): caches keyed on it
                        # (the generator's packed bf16 weights) see the overwrite
                        dist.broadcast(t.detach(), src=0, group=group)
        # t_cls_train.py:184-185: Adam, betas (0, 0.999), L2 weight decay lr/20 (not AdamW)
        if fused_adam is None:
            fused_adam = next(G.parameters()).is_cuda
        if fused_adam:  # same rule, one launch per model (optim.py / wu_adam_multi)
            from .optim import FusedAdam
            import functools
            Adam = functools.partial(FusedAdam, device_step=static_grads)
        else:
            Adam = torch.optim.Adam
        self.fused_adam = bool(fused_adam)
        self.g_opt = Adam(G.parameters(), lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
        self.d_opt = Adam(D.parameters(), lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
        if fused_adam and hasattr(G, "packed_weight_names"):
            # g_opt.step() (t_cls_train.py:273) must reach the bf16 operand copies the tensor-core
            # convolutions read: the same launch rewrites them
            self.g_opt.attach_packed(G)

    def _disc(self, x, c):
        if self.d_channels_last and not (self.d_autocast and x.dtype == torch.float32):
            x = x.contiguous(memory_format=torch.channels_last)  # (the stem kernels take NCHW fp32)
        if self.d_autocast and x.is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.D(x, c)[0].float()
        return self.D(x, c)[0]

    def _g_forward(self, images, c, masks):
        """The generator forward whose backward feeds the G update."""
        if not self._use_sink:
            return self.G(images, c, dropout_masks=masks)
        self.G._grad_sink = self.g_buckets
        try:
            return self.G(images, c, dropout_masks=masks)
        finally:
            self.G._grad_sink = None

    def _zero(self, opt, buckets):
        if buckets is not None:
            buckets.zero()
        else:
            opt.zero_grad(set_to_none=True)

    def step(self, images, c_real, c_target, masks_d=None, masks_g=None):
        """images (B,3,H,W) in [-1,1]; c_real = condition of the real images (one-hot label or
        estimator output); c_target = condition to transfer to.  Returns loss tensors (no sync)."""
        G, D = self.G, self.D
        # ---- discriminator update (t_cls_train.py:288-312)
        self._zero(self.d_opt, self.d_buckets)
        real = self._disc(images, c_real)
        if self.share_fake:
            shared = self._g_forward(images, c_target, masks_g)  # one forward for both updates
            fake_img = shared.detach()
        else:
            with torch.no_grad():  # the reference builds this graph and drops it with .detach() (:302-303)
                fake_img = G(images, c_target, dropout_masks=masks_d)
        fake = self._disc(fake_img, c_target)
        d_loss = dis_hinge(fake, real)
        d_loss.backward()
        # ---- generator update (t_cls_train.py:226-286)
        self._zero(self.g_opt, self.g_buckets)
        # The generator's forward does not read the discriminator, so it is enqueued BEFORE the
        # discriminator's gradient all-reduce is awaited and its Adam step applied: in data-parallel
        # runs the reduction (one bucket that completes at the very end of D's backward) hides under
        # the generator's convolutions.  Same arithmetic, same order of updates as the reference.
        fake_img = shared if self.share_fake else self._g_forward(images, c_target, masks_g)
        if self.d_buckets is not None:
            self.d_buckets.finish()
        self.d_opt.step()
        d_params = [p for p in D.parameters()]
        for p in d_params:  # D's weight gradients of this pass are discarded by the reference
            p.requires_grad_(False)
        try:
            fake = self._disc(fake_img, c_target)
            g_adv = gen_hinge(fake)
            if images.is_cuda and _ops().l1_per_sample_supported(fake_img, images):
                diff = _ops().l1_per_sample(fake_img, images)  # one pass each way (wu_l1_per_sample_*)
                g_l1 = diff.detach().mean()  # == F.l1_loss (equal sample sizes); logged only (:255)
            else:
                g_l1 = l1_loss(fake_img, images)  # logged only (:255)
                diff = (fake_img - images).abs().mean(dim=(1, 2, 3))
            lmda = (c_real - c_target).abs().mean(dim=1)
            loss_con = (diff / (lmda + self.eps_con)).mean()
            g_loss = g_adv + loss_con
            g_w = None
            if self.estimator is not None:
                g_w = F.mse_loss(self.estimator(fake_img), c_target)
                g_loss = g_loss + g_w
            g_loss.backward()
        finally:
            for p in d_params:
                p.requires_grad_(True)
        if self.g_buckets is not None:
            self.g_buckets.finish()
        self.g_opt.step()
        out = {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "g_loss_adv": g_adv.detach(),
               "g_loss_l1": g_l1.detach(), "loss_con": loss_con.detach()}
        if g_w is not None:
            out["g_loss_w"] = g_w.detach()
        return out


class GraphedGDStep:
    """The whole training iteration (D update + G update, both Adam steps, the gradient
    all-reduces when data parallel) captured ONCE in a CUDA graph and replayed per batch: the
    ~470 kernel launches of an iteration cost one graph launch on the host, and the GPU no longer
    idles between dependent kernels waiting for the next launch.

    What makes the iteration replayable: static input buffers; every gradient at a fixed address
    (GDTrainStep(static_grads=True)); the dropout draw counter and Adam's step count on the
    device (Conditional_UNet.use_device_dropout_counter, optim.FusedAdam(device_step=True)); TMA
    descriptors and table pointers baked at capture into memory the graph's pool keeps alive.
    Same arithmetic as GDTrainStep.step (which is what gets captured).

    `warmup` real training iterations run eagerly on the first batch before the capture (they are
    ordinary updates); the capture pass itself does not execute."""

    def __init__(self, trainer, images, c_real, c_target, warmup=3):
        if not (trainer.fused_adam and trainer.g_buckets is not None):
            raise RuntimeError("GraphedGDStep needs GDTrainStep(static_grads=True) (or data parallel) "
                               "with the fused optimiser")
        if getattr(trainer.G, "_drop_epoch", None) is None:
            trainer.G.use_device_dropout_counter(True)
        self.trainer = trainer
        self.inputs = tuple(torch.empty_like(t).copy_(t) for t in (images, c_real, c_target))
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                trainer.step(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(images.device)
        from . import _ops as K
        n0 = K.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = trainer.step(*self.inputs)
        self.library_launches = K.launch_count() - n0  # launches of this library per replay
        self._undo_capture_side_effects()

    def _undo_capture_side_effects(self):
        """The capture pass ran the host side of one iteration without executing it: the host mirrors
        of the step counts advanced by one although no update happened."""
        for opt in (self.trainer.g_opt, self.trainer.d_opt):
            for st in opt.state.values():
                if "step" in st:
                    st["step"] -= 1

    def step(self, images=None, c_real=None, c_target=None):
        """Copy the batch into the static buffers (skip by passing None and writing
        `self.inputs` yourself, e.g. straight from pinned host memory) and replay.  Returns the
        static loss tensors (overwritten by the next replay)."""
        if images is not None:
            for dst, src in zip(self.inputs, (images, c_real, c_target)):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for opt in (self.trainer.g_opt, self.trainer.d_opt):
            steps = [st["step"] for st in opt.state.values() if "step" in st]
            if steps:
                torch._foreach_add_(steps, 1)
        return self.outputs
