"""Drop-in for the reference's ops.py (ops.py:14-83): GAN / reconstruction losses and label
helpers.  These are scalar reductions over tiny tensors next to the generator, so they stay plain
PyTorch (SURVEY §8 f2 lists fused versions as a later row).  `from ops import *` in the trainers
also relies on the re-exported names F, np, torch, nn, Variable.
"""
import os  # noqa: F401

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Variable

xp = np


def _same_size(a, b):
    assert a.size() == b.size(), 'The size of a and b is different.{}!={}'.format(a.size(), b.size())


def soft_transform(x, std=0.05):
    """x + N(0, std) noise (ops.py:14-16)."""
    return x + torch.randn_like(x) * std


def adv_loss(a, b):
    _same_size(a, b)
    return F.mse_loss(a, b)


def l1_loss(a, b):
    _same_size(a, b)
    return F.l1_loss(a, b)


def feat_loss(a, b):
    """Mean of per-level L1 distances between two feature lists (ops.py:26-27)."""
    return torch.stack([F.l1_loss(u, v) for u, v in zip(a, b)]).mean()


def pred_loss(preds, labels, one_hot=False):
    """Weather-prediction loss: cross entropy on class indices when one_hot, else MSE
    (ops.py:29-40)."""
    if one_hot:
        return F.cross_entropy(preds, labels)
    return F.mse_loss(preds, labels)


def dis_hinge(dis_fake, dis_real):
    """Discriminator hinge loss (ops.py:42-45)."""
    return F.relu(1. - dis_real).mean() + F.relu(1. + dis_fake).mean()


def gen_hinge(dis_fake):
    """Generator hinge loss (ops.py:47-48)."""
    return (-dis_fake).mean()


def vector_to_one_hot(vec):
    """One-hot of the arg-max along dim 0 (ops.py:50-54)."""
    out = torch.zeros_like(vec)
    out.scatter_(0, torch.argmax(vec, 0, keepdim=True), 1)
    return out


def get_rand_labels(num_classes, batch_size, one_hot=False):
    """Uniform(-1, 1) labels on the GPU (ops.py:56-60).  The reference's one_hot branch calls
    F.one_hot on a float tensor and raises; the same call is kept so the failure mode matches."""
    label = torch.empty(batch_size, num_classes).uniform_(-1, 1)
    if one_hot:
        label = F.one_hot(label, num_classes)
    return label.to('cuda')


def get_sequential_labels(num_classes, batch_size, one_hot=False):
    """0,1,..,nc-1,0,1,.. truncated to batch_size; one-hot rows when asked (ops.py:62-71)."""
    idx = torch.arange(batch_size) % num_classes
    if one_hot:
        return torch.eye(num_classes, dtype=torch.float32)[idx].to('cuda')
    return idx.to(torch.float32).to('cuda')


def Variable_Float(x, batch_size):
    """(batch_size, 1) CUDA float tensor filled with x (ops.py:73-74)."""
    return Variable(torch.full((batch_size, 1), float(x), device='cuda'), requires_grad=False)


def make_table_img(images, ref_images, results):
    """Stack the inputs and the per-reference results along H (ops.py:77-83)."""
    return torch.cat([images] + results, dim=2)
