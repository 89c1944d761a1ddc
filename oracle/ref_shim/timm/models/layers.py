"""Stand-ins for the three timm symbols imported by the reference at
models/hit_sir_pro.py:6.  Semantics follow timm 1.0.15 (requirements.txt:50):
  * to_2tuple: int -> (int, int), iterables pass through;
  * trunc_normal_: torch.nn.init.trunc_normal_ (same algorithm timm vendors);
  * DropPath: only instantiated when drop_path > 0 (hit_sir_pro.py:658); identity in eval.
Test infrastructure only - never imported by the product package."""
import collections.abc
from itertools import repeat

import torch
from torch import nn


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0., scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask
