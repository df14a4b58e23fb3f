"""Script-compatibility shims: what the reference's trainer / inference scripts need
in order to run on torch 2.x and under ``torchrun`` with this package in place of cunet.py /
disc.py — without editing their training logic.

The reference was written for torch 1.1 (Pipfile:10-11).  Three idioms no longer work:
  * ``test_data_iter.next()`` (t_cls_train.py:218, t_est_train.py:196): DataLoader iterators lost
    ``.next`` → `install_torch2_patches()` restores it;
  * ``torch.load(args.estimator_path)`` of a PICKLED MODULE (t_cls_train.py:172, demo.py:56,
    inference/inf_1year_signals.py:87): torch >= 2.6 defaults to ``weights_only=True`` →
    `load_module()` / the patch's default;
  * one process, one GPU, ``shuffle=True`` loaders (t_cls_train.py:187-216): under torchrun every rank
    would see the same batches → `make_loader()` shards with a DistributedSampler (and keeps
    ``ImbalancedDatasetSampler``-style samplers by sharding their index stream), `set_epoch()`
    reshuffles per epoch.
`train_epochs()` is the loop of t_cls_train.py:387-437 / t_est_train.py:353-407 (zip of the two
loaders, label preparation, D update + G update, checkpoint every `save_per_step`) driving
train_step.GDTrainStep; data sets, augmentation, TensorBoard and CLI stay the scripts' own.
"""
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Sampler
from torch.utils.data.distributed import DistributedSampler


def install_torch2_patches():
    """Make the unmodified reference scripts importable / runnable on torch 2.x:
    ``iterator.next()`` on DataLoader iterators, and ``torch.load`` defaulting to
    ``weights_only=False`` (the scripts load pickled nn.Modules they wrote themselves)."""
    from torch.utils.data import dataloader as _dl
    if not hasattr(_dl._BaseDataLoaderIter, "next"):
        _dl._BaseDataLoaderIter.next = _dl._BaseDataLoaderIter.__next__
    if not getattr(torch.load, "_wu_patched", False):
        _orig = torch.load

        def load(*args, **kwargs):
            kwargs.setdefault("weights_only", False)
            return _orig(*args, **kwargs)
        load._wu_patched = True
        load._wu_orig = _orig
        torch.load = load


def load_module(path, map_location="cpu"):
    """``torch.load`` of a pickled module (estimator / classifier checkpoints,
    t_cls_train.py:172-177): explicit ``weights_only=False``; the file must be trusted."""
    fn = getattr(torch.load, "_wu_orig", torch.load)
    return fn(path, map_location=map_location, weights_only=False)


def init_distributed(backend="nccl"):
    """One process per GPU under torchrun: reads RANK / LOCAL_RANK / WORLD_SIZE, selects the GPU and
    initialises the process group.  Returns (rank, world, device).  A plain ``python script.py`` run
    gives (0, 1, cuda:0 or cpu) and no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl" and torch.cuda.is_available():
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
    else:
        device = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if device.type == "cuda" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, device


class ShardedSampler(Sampler):
    """Rank's share of another sampler's index stream (e.g. the reference's
    ImbalancedDatasetSampler, t_cls_train.py:196-202): every rank draws the same stream (same seed)
    and keeps indices rank, rank + world, ...; truncated so that all ranks get the same count."""

    def __init__(self, sampler, rank, world, seed=0):
        self.sampler, self.rank, self.world, self.seed, self.epoch = sampler, rank, world, seed, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __iter__(self):
        state = torch.random.get_rng_state()
        torch.manual_seed(self.seed + self.epoch)  # identical stream on every rank
        try:
            idx = list(iter(self.sampler))
        finally:
            torch.random.set_rng_state(state)
        n = len(idx) // self.world * self.world
        return iter(idx[self.rank:n:self.world])

    def __len__(self):
        return len(self.sampler) // self.world


def make_loader(dataset, batch_size, shuffle=True, sampler=None, num_workers=0, drop_last=True,
                rank=None, world=None, seed=0, **kw):
    """The DataLoaders of t_cls_train.py:187-216, rank-aware: with world > 1 the data set (or the
    given sampler's index stream) is sharded across ranks; `batch_size` is PER RANK, as in the
    reference's single-GPU runs (weak scaling)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world > 1:
        if sampler is not None:
            sampler = ShardedSampler(sampler, rank, world, seed)
        else:
            sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=shuffle,
                                         seed=seed, drop_last=drop_last)
        shuffle = False
    elif sampler is not None:
        shuffle = False
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, sampler=sampler,
                      num_workers=num_workers, drop_last=drop_last, pin_memory=torch.cuda.is_available(),
                      **kw)


def set_epoch(loader, epoch):
    """Reshuffle a rank-aware loader for a new epoch (no-op for plain loaders)."""
    s = getattr(loader, "sampler", None)
    if hasattr(s, "set_epoch"):
        s.set_epoch(epoch)


def first_batches(loader, n, device):
    """``[tuple(d.to('cuda') for d in it.next()) for i in range(n)]`` (t_cls_train.py:217-219)."""
    it = iter(loader)
    return [tuple(d.to(device) for d in next(it)) for _ in range(n)]


def one_hot(labels, num_classes, device):
    """``torch.eye(5)[c].to('cuda')`` (t_cls_train.py:421-422) without the host round trip."""
    return torch.nn.functional.one_hot(labels.to(device).long(), num_classes).float()


def train_epochs(step, train_loader, random_loader, num_classes, device, epochs=1, supervised=True,
                 estimator=None, batch_size=None, save_per_step=0, on_save=None, start_step=0,
                 on_log=None):
    """The iteration loop of t_cls_train.py:387-437: zip of the two loaders; incomplete batches
    skipped (:413-414, :424-425); supervised: one-hot labels (:421-422), otherwise the target
    condition is the frozen estimator's prediction for the random images (:424); then D update and
    G update (`step` = train_step.GDTrainStep.step or GraphedGDStep.step).  Returns the global step."""
    global_step = start_step
    for epoch in range(epochs):
        set_epoch(train_loader, epoch)
        set_epoch(random_loader, epoch)
        for data, rand_data in zip(train_loader, random_loader):
            global_step += 1
            if save_per_step and on_save is not None and global_step % save_per_step == 0:
                on_save(epoch, global_step)
            images, c_d = (d.to(device, non_blocking=True) for d in data[:2])
            rand_images, c_r = (d.to(device, non_blocking=True) for d in rand_data[:2])
            if batch_size is not None and images.size(0) != batch_size:
                continue
            if supervised:
                c_target = one_hot(c_r, num_classes, device)
                c_real = one_hot(c_d, num_classes, device)
            else:
                with torch.no_grad():
                    c_target = estimator(rand_images).detach()
                c_real = c_d.float() if c_d.dim() == 2 else one_hot(c_d, num_classes, device)
            losses = step(images, c_real, c_target)
            if on_log is not None:
                on_log(global_step, losses)
    return global_step
