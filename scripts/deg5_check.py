"""Localise the known gradient deviation on graphs with in-degree > 4.

    python scripts/deg5_check.py            # on a GPU box

Runs st_gat_3 (full width) on two random trees with up to 4 children per node and prints, for every parameter, the
error of its gradient against the CPU oracle — with the head-averaged output layer evaluated aggregate-first
(gat_wide.cu) and projection-first (gat_layer.cu), and for a control batch of bifurcating trees of the same size that
takes the same chunk kernels (more than 384 nodes per tree switches the per-tree kernels off).  Whichever column
turns bad names the kernel family whose general-degree branch is wrong.
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from helpers import FULL_MODELS
from oracle import dgl_ops, models as om, pe as ope
from spgnn_b200 import graph as sg, models as sm, ops, pe as spe, synth, stack
mods = dict(dgl_ops=dgl_ops, om=om, ope=ope, sg=sg, sm=sm, ops=ops, spe=spe, synth=synth)
kind, cfg = FULL_MODELS["st_gat_3"]
cw = torch.tensor([0.2] + [0.8] * 21)


def batch(max_children, sizes, seed=7):
    rng = np.random.default_rng(seed)
    scans = []
    for n in sizes:
        a = T._random_tree_adj(n, max_children, rng)
        scans.append(dict(adj=a, fvs=np.maximum(rng.standard_normal((n, 1024)), 0).astype(np.float32),
                          fvs_out=rng.standard_normal((n, 22)).astype(np.float32),
                          labels=rng.integers(0, 22, n).astype(np.int64)))
    return scans


def grads(scans, wide):
    stack.WIDE_OUTPUT_LAYER = wide
    torch.manual_seed(0)
    onet = om.GNNNet(kind, cfg); onet.init_like_reference(); onet.eval()
    net = sm.GATNet(**cfg).cuda(); net.load_state_dict(onet.state_dict(), strict=True); net.eval()
    og, g = T._oracle_batch(mods, scans), T._device_batch(mods, scans)
    y = torch.from_numpy(np.concatenate([s["labels"] for s in scans]))
    mask = torch.ones(y.numel(), dtype=torch.bool)
    om.cross_entropy_masked(onet(og)[0], y, mask, cw).backward()
    ops.masked_cross_entropy(net(g)[0], y.cuda(), cw.cuda(), mask=mask.cuda()).backward()
    ref = dict(onet.named_parameters())
    return g.max_degree(), {k: float((p.grad.cpu().double() - ref[k].grad.double()).abs().max() / ref[k].grad.abs().max())
                            for k, p in net.named_parameters() if ref[k].grad is not None}


cols = {}
for name, mc, sizes in (("deg<=6", 4, (150, 90)), ("control deg<=4, chunk kernels", 2, (500, 90))):
    for wide in (True, False):
        md, e = grads(batch(mc, sizes), wide)
        cols[f"{name} | max_degree {md} | wide={wide}"] = e
keys = list(next(iter(cols.values())))
for c in cols:
    print("column:", c)
for k in keys:
    print(f"{k:42s} " + "  ".join(f"{cols[c][k]:9.2e}" for c in cols))
