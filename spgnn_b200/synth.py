"""Seeded synthetic airway-tree batches (host side, numpy).

Input generator for tests and bench.  The reference has no
generator of its own; the shapes follow what its stage-1 pickles hold
(/root/reference/job_runner.py:796-805: ``fvs`` [n,1024], ``adj`` uint8 [n,n] =
tree ∪ I, ``labels`` [n], ``fvs_out`` [n,22]) and the invariants its runtime
asserts need (job_runner.py:1741 — 21 distinct anchors ⇒ n ≥ 21;
dataset.py:418-419 — adj symmetric with unit diagonal; parent index < child
index, job_runner.py:1713).

Randomness is a hand-rolled Philox4x32-10 so that the *integer* part (tree
shape, labels) has a bit-exact device twin in ``csrc/synth.cu``; the float part
(Box-Muller normals) agrees with the device twin to rounding only.

Counter layout: ``(i0, i1, stream, tree)``; key = ``(seed, 0x5350474e)``.
  stream 0: bifurcation picks (i0 = step)      stream 1: k_t for ragged mode
  stream 2: label placement (i0 = draw)         stream 3: fvs normals (i0 = elem/4)
  stream 4: fvs_out normals (i0 = elem/4)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
KEY1 = 0x5350474E  # "SPGN"

NR_CLASS = 22
FV_DIM = 1024


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 over broadcastable uint32 arrays → four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _mulhi_pick(r, n):
    """Uniform integer in [0, n) from a uint32 draw: (r * n) >> 32 (exact)."""
    return ((r.astype(np.uint64) * np.asarray(n, dtype=np.uint64)) >> np.uint64(32)).astype(np.int64)


def _normals(n_elem, stream, tree, seed):
    """``n_elem`` standard normals for one tree (Box-Muller on Philox words)."""
    nblk = (n_elem + 3) // 4
    i0 = np.arange(nblk, dtype=np.uint32)
    w = philox4x32(i0, np.uint32(0), np.uint32(stream), np.uint32(tree), seed, KEY1)
    u = [(x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0) for x in w]
    r0 = np.sqrt(-2.0 * np.log(u[0]))
    r1 = np.sqrt(-2.0 * np.log(u[2]))
    t0 = 2.0 * np.pi * u[1]
    t1 = 2.0 * np.pi * u[3]
    out = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=1)
    return out.reshape(-1)[:n_elem].astype(np.float32)


def tree_sizes(tree_ids, seed=1234, ragged=False, k_fixed=150, k_lo=120, k_hi=180):
    """Bifurcation count k_t per tree (n_t = 2 k_t + 1)."""
    tree_ids = np.asarray(tree_ids, dtype=np.uint32)
    if not ragged:
        return np.full(tree_ids.shape, k_fixed, dtype=np.int64)
    r = philox4x32(np.uint32(0), np.uint32(0), np.uint32(1), tree_ids, seed, KEY1)[0]
    return k_lo + _mulhi_pick(r, k_hi - k_lo + 1)


def tree_parents(tree_id, k, seed=1234):
    """Parent array of one random full-bifurcation tree (root 0 has parent -1).

    Start with root 0; at step j pick a uniformly random current leaf and give it
    children 2j+1, 2j+2 (creation order ⇒ parent < child).
    """
    r = philox4x32(np.arange(k, dtype=np.uint32), np.uint32(0), np.uint32(0), np.uint32(tree_id), seed, KEY1)[0]
    parent = np.full(2 * k + 1, -1, dtype=np.int64)
    leaves = np.zeros(k + 1, dtype=np.int64)
    for j in range(k):
        p = int((int(r[j]) * (j + 1)) >> 32)
        node = leaves[p]
        parent[2 * j + 1] = node
        parent[2 * j + 2] = node
        leaves[p] = 2 * j + 1
        leaves[j + 1] = 2 * j + 2
    return parent


def tree_labels(tree_id, n, seed=1234):
    """Labels 1..21 on 21 distinct nodes (partial Fisher-Yates on Philox draws), 0 elsewhere."""
    r = philox4x32(np.arange(NR_CLASS - 1, dtype=np.uint32), np.uint32(0), np.uint32(2), np.uint32(tree_id), seed, KEY1)[0]
    perm = np.arange(n, dtype=np.int64)
    y = np.zeros(n, dtype=np.int64)
    for j in range(min(NR_CLASS - 1, n)):      # trees smaller than 21 nodes (tests only) get labels 1..n
        p = j + int((int(r[j]) * (n - j)) >> 32)
        perm[j], perm[p] = perm[p], perm[j]
        y[perm[j]] = j + 1
    return y


def adj_from_parents(parent):
    """Dense uint8 adjacency = symmetric tree ∪ identity (dataset.py:418-419 form)."""
    n = parent.shape[0]
    adj = np.eye(n, dtype=np.uint8)
    c = np.arange(1, n)
    adj[c, parent[1:]] = 1
    adj[parent[1:], c] = 1
    return adj


@dataclass
class Scan:
    """One synthetic 'scan' in the stage-1 pickle layout (job_runner.py:796-805)."""
    adj: np.ndarray      # uint8 [n,n]
    fvs: np.ndarray      # float32 [n,1024]
    fvs_out: np.ndarray  # float32 [n,22]
    labels: np.ndarray   # int64 [n]
    parent: np.ndarray   # int64 [n]
    tree_id: int


def make_scan(tree_id, seed=1234, ragged=False, fv_dim=FV_DIM, k=None, features=True):
    if k is None:
        k = int(tree_sizes([tree_id], seed, ragged)[0])
    parent = tree_parents(tree_id, k, seed)
    n = parent.shape[0]
    adj = adj_from_parents(parent)
    labels = tree_labels(tree_id, n, seed)
    if features:
        fvs = np.maximum(_normals(n * fv_dim, 3, tree_id, seed), 0.0).reshape(n, fv_dim)
        fvs_out = (3.0 * _normals(n * NR_CLASS, 4, tree_id, seed)).astype(np.float32).reshape(n, NR_CLASS)
    else:
        fvs = np.zeros((n, fv_dim), np.float32)
        fvs_out = np.zeros((n, NR_CLASS), np.float32)
    return Scan(adj, fvs, fvs_out, labels, parent, int(tree_id))


def make_scans(first_tree, count, **kw):
    return [make_scan(first_tree + i, **kw) for i in range(count)]
