"""Data-parallel host logic on CPU (gloo, world_size 2): the gradient buckets all-reduce to the
mean of the per-rank gradients, parameters without a gradient are tolerated, replicas start
identical, and a 2-rank G+D iteration equals the 1-rank iteration on the concatenated batch.
The generator here is a small PyTorch stand-in with the Conditional_UNet call signature (the real
one needs a GPU); the discriminator and the step logic are the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class TinyG(nn.Module):
    """Per-sample generator stand-in: forward(x, c, dropout_masks=None) -> image."""

    def __init__(self, nc=5):
        super().__init__()
        self.conv = nn.Conv2d(3, 3, 3, padding=1)
        self.l1 = nn.Linear(nc, 3)
        self.unused = nn.Embedding(nc, nc)  # never receives a gradient (like adain*.emb)

    def forward(self, x, c, dropout_masks=None):
        return torch.tanh(self.conv(x) + self.l1(c)[:, :, None, None])


def _make(seed):
    from weather_unet_b200.disc import SNDisc
    torch.manual_seed(seed)
    return TinyG(), SNDisc(5)


def _data(n):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(n, 3, 32, 32, generator=g) * 2 - 1
    cr = torch.eye(5)[torch.randint(0, 5, (n,), generator=g)]
    ct = torch.eye(5)[torch.randint(0, 5, (n,), generator=g)]
    return x, cr, ct


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from weather_unet_b200.train_step import GDTrainStep, GradBuckets
    try:
        # ---- bucket mechanics on a plain module
        torch.manual_seed(0)
        m = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 2))
        ref = [p.detach().clone() for p in m.parameters()]
        buckets = GradBuckets(m.named_parameters(), bucket_bytes=64)
        assert len(buckets.buckets) > 1
        buckets.attach_autograd_hooks()
        xs = torch.arange(16, dtype=torch.float32).view(4, 4) / 10
        buckets.zero()
        m(xs[2 * rank:2 * rank + 2]).pow(2).mean().backward()
        buckets.finish()
        m2 = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 2))
        for p, r in zip(m2.parameters(), ref):
            p.data.copy_(r)
        m2(xs).pow(2).mean().backward()
        for p, q in zip(m.parameters(), m2.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), "bucket all-reduce != full-batch gradient"
        # ---- full iteration: rank r starts from a different seed; broadcast makes them identical
        G, D = _make(seed=10 + rank)
        step = GDTrainStep(G, D, lr=1e-3, d_autocast=False)
        x, cr, ct = _data(4)
        sl = slice(2 * rank, 2 * rank + 2)
        for _ in range(2):
            losses = step.step(x[sl], cr[sl], ct[sl])
        if rank == 0:
            ret["params"] = {k: v.detach().clone() for k, v in list(G.state_dict().items()) +
                             [("D." + k, v) for k, v in D.state_dict().items()]}
            ret["d_loss"] = float(losses["d_loss"])
        gsum = [p.detach().clone() for p in G.parameters()]
        for t in gsum:
            dist.all_reduce(t)
        for p, t in zip(G.parameters(), gsum):
            assert torch.allclose(p.detach() * world, t, atol=1e-6), "replicas diverged"
        ret[f"ok{rank}"] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_step_matches_single_rank():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret.get("ok0") and ret.get("ok1")
    # single process, same initial weights as rank 0, global batch of 4
    from weather_unet_b200.train_step import GDTrainStep
    G, D = _make(seed=10)
    step = GDTrainStep(G, D, lr=1e-3, d_autocast=False)
    x, cr, ct = _data(4)
    for _ in range(2):
        step.step(x, cr, ct)
    single = {k: v for k, v in list(G.state_dict().items()) + [("D." + k, v) for k, v in D.state_dict().items()]}
    for k, v in ret["params"].items():
        if k.endswith("_u") or k.endswith("_v"):
            continue  # power-iteration buffers see different forward counts per shard: not compared
        assert torch.allclose(v, single[k], atol=2e-4, rtol=1e-3), f"{k}: DP(2) != single-rank"


def test_static_grads_single_process_equal_plain_step():
    """GDTrainStep(static_grads=True) — the bucket bookkeeping without a process group, what the
    CUDA-graph replay relies on — gives the same parameters as the plain step, keeps every .grad at a
    fixed address across iterations, and leaves the never-used parameter without a gradient."""
    from weather_unet_b200.train_step import GDTrainStep
    x, cr, ct = _data(4)
    Ga, Da = _make(3)
    Gb, Db = _make(3)
    ta = GDTrainStep(Ga, Da, lr=1e-3, d_autocast=False, fused_adam=False)
    tb = GDTrainStep(Gb, Db, lr=1e-3, d_autocast=False, fused_adam=False, static_grads=True)
    assert tb.g_buckets is not None and not tb.g_buckets.collective and not tb.distributed
    ptrs = None
    for it in range(3):
        la = ta.step(x, cr, ct)
        lb = tb.step(x, cr, ct)
        assert all(torch.equal(la[k], lb[k]) for k in la), it
        now = {n: p.grad.data_ptr() for n, p in list(Gb.named_parameters()) + list(Db.named_parameters())
               if p.grad is not None}
        if ptrs is not None:
            assert now == ptrs, "a gradient moved between iterations"
        ptrs = now
    assert Gb.unused.weight.grad is None
    for (n, p), (_, q) in zip(list(Ga.named_parameters()) + list(Da.named_parameters()),
                              list(Gb.named_parameters()) + list(Db.named_parameters())):
        assert torch.equal(p, q), n


def test_packed_weight_cache_protocol():
    """_generator.PackedWeights: one persistent buffer pair per weight (fixed addresses), repack when
    the master's version counter or storage changes, no repack after a writer called mark_fresh
    (what optim.FusedAdam does after rewriting the copies itself), repack after invalidate()."""
    from weather_unet_b200 import _generator as gen
    calls = []
    orig = gen.K.pack_conv3x3_weights_into
    gen.K.pack_conv3x3_weights_into = lambda w, wf, wd: calls.append(w._version)
    try:
        pw = gen.PackedWeights()
        w = torch.nn.Parameter(torch.randn(64, 128, 3, 3))
        wf, wd = pw.get("a", w)
        assert wf.shape == (64, 9 * 128) and wd.shape == (128, 9 * 64) and len(calls) == 1
        assert pw.get("a", w)[0] is wf and len(calls) == 1            # unchanged master: cache hit
        with torch.no_grad():
            w.add_(1.0)                                               # torch optimiser step
        assert pw.get("a", w)[0] is wf and len(calls) == 2            # same buffers, repacked
        torch.autograd.graph.increment_version(w)                     # raw-pointer writer ...
        pw.mark_fresh("a", w)                                         # ... that rewrote the copies itself
        pw.get("a", w)
        assert len(calls) == 2
        torch.autograd.graph.increment_version(w)                     # raw-pointer writer that did not
        pw.get("a", w)
        assert len(calls) == 3
        pw.invalidate()
        assert pw.get("a", w)[1] is wd and len(calls) == 4
        w2 = torch.nn.Parameter(torch.randn(64, 64, 3, 3))            # another shape under the same name
        assert pw.get("a", w2)[0].shape == (64, 9 * 64) and len(calls) == 5
    finally:
        gen.K.pack_conv3x3_weights_into = orig
