"""Replay of the device's non-smooth decisions inside the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

The hot path holds three kinds of non-smooth operations: LeakyReLU / ReLU kinks (GATConv's attention logits,
GIN's MLP, SAGEConv's pool projection), SAGEConv's element-wise max over neighbours, and dropout.  Their GRADIENTS
are discontinuous functions of the forward values: a logit within rounding distance of 0 takes slope 1 in one
implementation and 0.2 in the other, two nearly tied neighbours route a gradient row to different sources, and two
RNG streams draw different masks.  The forward values agree to 1e-5 either way; the gradients do not.

A :class:`Tape` holds, in call order, the decisions the DEVICE took (``spgnn_b200.ops.KINK_TRACE``): the sign of
every pre-activation, the arg-max source of every pooled element, every dropout mask.  While a tape is active
(``with kinks.use(tape):``) the oracle's ops take the same branch instead of deciding for themselves, so the
comparison of gradients is exact, in train mode too.  Every replayed decision is also checked against the
oracle's own: a disagreement is only accepted when the value is within ``tol`` of the kink / the tie (relative to
the tensor's largest magnitude) — anything else is reported in ``tape.violations`` and fails the test.
Without an active tape every function here is the plain torch op.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_active = None


class Tape:
    def __init__(self, records, tol=1e-4):
        self.records = [(k, v.detach().cpu() if torch.is_tensor(v) else v) for k, v in records]
        self.pos, self.tol = 0, tol
        self.flips = 0            # decisions where device and oracle differ (inside the tolerance)
        self.decisions = 0
        self.violations = []      # (record index, kind, size of the disagreement relative to the tensor scale)

    def take(self, kind):
        if self.pos >= len(self.records):
            raise AssertionError(f"kink tape exhausted: the oracle asked for a {kind!r} record the device never wrote")
        k, v = self.records[self.pos]
        if k != kind:
            raise AssertionError(f"kink tape out of step at record {self.pos}: device wrote {k!r}, oracle wants {kind!r}")
        self.pos += 1
        return v

    def done(self):
        return self.pos == len(self.records)


@contextlib.contextmanager
def use(tape):
    global _active
    prev, _active = _active, tape
    try:
        yield tape
    finally:
        _active = prev


def _sign_replay(x, slope):
    t = _active
    m = t.take("sign").reshape(x.shape)
    own = x.detach() > 0
    diff = own != m
    t.decisions += m.numel()
    n = int(diff.sum())
    if n:
        t.flips += n
        worst = float(x.detach().abs()[diff].max() / x.detach().abs().max().clamp(min=1e-30))
        if worst > t.tol:
            t.violations.append((t.pos - 1, "sign", worst))
    return torch.where(m, x, x * slope)


def leaky_relu(x, slope=0.01):
    if _active is None:
        return F.leaky_relu(x, slope)
    return _sign_replay(x, slope)


def relu(x):
    if _active is None:
        return F.relu(x)
    return _sign_replay(x, 0.0)


def seg_max_nodes(m, src, dst, n):
    """max over in-neighbours of node rows m[src] → [n, D]; with a tape: the device's arg-max source per element."""
    if _active is None:
        out = torch.full((n, m.shape[1]), -math.inf, dtype=m.dtype)
        out = out.scatter_reduce(0, dst.view(-1, 1).expand(-1, m.shape[1]), m[src], reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)      # DGL zero-fills empty rows
    t = _active
    arg = t.take("argmax").long()                                            # [n, D] source node ids (-1: no in-edge)
    ref = torch.full((n, m.shape[1]), -math.inf, dtype=m.dtype)
    ref = ref.scatter_reduce(0, dst.view(-1, 1).expand(-1, m.shape[1]), m.detach()[src], reduce="amax", include_self=True)
    has = arg >= 0
    out = torch.where(has, m.gather(0, arg.clamp(min=0)), torch.zeros_like(ref))
    gap = torch.where(has, ref - out.detach(), torch.zeros_like(ref))
    t.decisions += arg.numel()
    n_diff = int((gap > 0).sum())
    if n_diff:
        t.flips += n_diff
        worst = float(gap.max() / m.detach().abs().max().clamp(min=1e-30))
        if worst > t.tol:
            t.violations.append((t.pos - 1, "argmax", worst))
    return out


def dropout(module, x):
    """``module`` is the nn.Dropout of the oracle layer; with a tape and in train mode the device's mask (already
    scaled by 1/(1-p)) replaces torch's draw."""
    if _active is None or not module.training or module.p == 0.0:
        return module(x)
    mask = _active.take("drop").to(x.dtype).reshape(x.shape)
    return x * mask


class LeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.01):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return leaky_relu(x, self.negative_slope)


class Dropout(nn.Dropout):
    def forward(self, x):
        if _active is None or not self.training or self.p == 0.0:
            return super().forward(x)
        return x * _active.take("drop").to(x.dtype).reshape(x.shape)
