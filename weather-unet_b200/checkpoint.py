"""Checkpoint files of the reference trainers.

The trainers write ``{'inference': G.state_dict(), 'discriminator': D.state_dict(), 'epoch': int,
'global_step': int}`` with ``torch.save`` (t_cls_train.py:399-406, t_est_train.py:365-373) and resume /
infer from them with ``load_state_dict(sd['inference'])`` (t_cls_train.py:158-166, demo.py:52-53,
inference/inf_*.py).  The generator and discriminator of this package keep the reference's state_dict
keys, so the files are interchangeable in both directions; these helpers only add what torch 2.x needs
(``weights_only`` handling, tensors moved to the CPU before saving, rank-0-only writes under torchrun).
"""
import os

import torch
import torch.distributed as dist


def checkpoint_name(save_dir, name, epoch, global_step):
    """`<save_dir>/<name>/<name>_e{epoch:04d}_s{step}.pt` — the reference's naming rule
    (t_cls_train.py:399)."""
    return os.path.join(save_dir, name, f"{name}_e{epoch:04d}_s{global_step}.pt")


def save_checkpoint(path, generator, discriminator, epoch, global_step):
    """Write a reference-format checkpoint (rank 0 only when a process group is initialised)."""
    if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
        return None
    sd = {"inference": {k: v.detach().cpu() for k, v in generator.state_dict().items()},
          "discriminator": {k: v.detach().cpu() for k, v in discriminator.state_dict().items()},
          "epoch": int(epoch), "global_step": int(global_step)}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(sd, tmp)
    os.replace(tmp, path)  # never leave a half-written file behind
    return path


def load_checkpoint(path, generator=None, discriminator=None, map_location="cpu"):
    """Read a checkpoint written by the reference trainers or by `save_checkpoint`.  The file holds
    only tensors and ints, so it is read with ``weights_only=True``.  When modules are given their
    weights are loaded strictly, like the reference does (missing / unexpected keys raise).
    Returns (epoch, global_step); files without those fields (inference-only exports) give (0, 0)."""
    sd = torch.load(path, map_location=map_location, weights_only=True)
    if "inference" not in sd:
        raise KeyError(f"{path}: no 'inference' entry; keys are {sorted(sd)}")
    if generator is not None:
        generator.load_state_dict(sd["inference"], strict=True)
    if discriminator is not None:
        if "discriminator" not in sd:
            raise KeyError(f"{path}: no 'discriminator' entry")
        discriminator.load_state_dict(sd["discriminator"], strict=True)
    return int(sd.get("epoch", 0)), int(sd.get("global_step", 0))


def latest_checkpoint(save_dir, name):
    """The newest `<name>_e*_s*.pt` under `<save_dir>/<name>` (the reference resumes from
    ``sorted(glob(...))[-1]``, t_cls_train.py:158-161), or None."""
    d = os.path.join(save_dir, name)
    if not os.path.isdir(d):
        return None
    files = sorted(f for f in os.listdir(d) if f.startswith(name + "_e") and f.endswith(".pt"))
    return os.path.join(d, files[-1]) if files else None
