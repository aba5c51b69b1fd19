"""The reference's dormant positional-encoding extras, on the device:

  compute_eigen_basis  job_runner.py:1630-1645  Laplacian-eigenvector PE (``g.ndata['eigvec']``)
  laplacian_pos_loss   job_runner.py:1803-1825  trace(pᵀ L p) + λ‖pᵀp − I‖_F per tree
  dist_pos_loss        job_runner.py:1827-1861  affinity-vs-hop-distance smooth-L1 + label-anchor compactness

None of them is reached by a shipped exp_settings file (``USE_DIST_LOSS = False`` everywhere, the eigenvector PE call
is commented out), so they are not kernels of this library: they are written with device tensor algebra over the
batch's CSR (dense per-tree blocks of at most a few hundred nodes; ``torch.linalg.eigh`` for the symmetric
normalised Laplacian) and checked against the literal restatement in ``oracle/pe.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ._lib import SpgnnError


def _tree_blocks(g):
    no = g.node_off.tolist()
    return [(no[i], no[i + 1]) for i in range(g.batch_size)]


def _norm_laplacian(g, lo, hi):
    """I − D^-1/2 A D^-1/2 of one tree of the batch, A as DGL's adjacency_matrix (A[src, dst], self loops included if
    the graph has them), D = in-degrees clipped at 1 (job_runner.py:1632-1634, :1816-1818)."""
    n = hi - lo
    m = (g.src >= lo) & (g.src < hi)
    a = torch.zeros(n, n, dtype=torch.float64, device=g.device)
    a[g.src[m] - lo, g.dst[m] - lo] = 1.0
    deg = g.in_degrees()[lo:hi].clamp(min=1).to(torch.float64)
    d = deg.pow(-0.5)
    return torch.eye(n, dtype=torch.float64, device=g.device) - d[:, None] * a * d[None, :]


def compute_eigen_basis(g, pos_enc_dim=39, store=True):
    """``eigvec`` [N, pos_enc_dim]: per tree the eigenvectors 1..pos_enc_dim of the normalised Laplacian in increasing
    eigenvalue order, zero-padded when the tree has fewer nodes.  (Eigenvectors are defined up to sign / rotation
    inside degenerate eigenspaces; the reference's ``np.linalg.eig`` pick is as arbitrary as this one.)"""
    out = torch.zeros(g.num_nodes, pos_enc_dim, dtype=torch.float32, device=g.device)
    for lo, hi in _tree_blocks(g):
        L = _norm_laplacian(g, lo, hi)
        if not torch.allclose(L, L.t()):
            raise SpgnnError("compute_eigen_basis: the adjacency must be symmetric")
        val, vec = torch.linalg.eigh(L)                       # ascending
        k = min(pos_enc_dim, hi - lo - 1)
        out[lo:hi, :k] = vec[:, 1:1 + k].float()
    if store:
        g.ndata["eigvec"] = out
    return out


def laplacian_pos_loss(g, p=None, lamb=0.1, pos_enc_dim=39):
    p = g.ndata["p"] if p is None else p
    losses = []
    for lo, hi in _tree_blocks(g):
        pb = p[lo:hi]
        n = hi - lo
        pz = pb - pb.mean(0, keepdim=True).detach()
        pn = pz / (pb.std(0, keepdim=True) + 1e-7).detach()
        L = _norm_laplacian(g, lo, hi).float()
        loss1 = torch.trace(pn.t() @ L @ pn)
        ptp = pn.t() @ pn - torch.eye(pn.shape[1], device=p.device)
        losses.append((loss1 + lamb * torch.linalg.norm(ptp, "fro")) / (pos_enc_dim * n))
    return torch.stack(losses).mean()


class DistPosLoss:
    """``dist_pos_loss`` with its running state (``cached_mean_pos_enc``, exponential average 0.15 / 0.85)."""

    def __init__(self, nr_class=22, pos_enc_dim=39):
        self.nr_class, self.pos_enc_dim = nr_class, pos_enc_dim
        self.cached_mean_pos_enc = None

    def __call__(self, g, all_pos_encs_cache, p=None, y=None, batch_stats_init=None):
        """``all_pos_encs_cache[i]``: the [n_i, n_i] hop-distance / diameter matrix of tree i (second output of
        ``generate_distant_pos_enc``).  ``batch_stats_init`` [B, C-1, pos_enc_dim] replaces the reference's
        ``torch.rand`` fill of the entries of absent labels (pass it for reproducible comparisons)."""
        p = g.ndata["p"] if p is None else p
        y = g.ndata["y"] if y is None else y
        B, C = g.batch_size, self.nr_class
        stats = torch.rand(B, C - 1, self.pos_enc_dim, device=p.device) if batch_stats_init is None \
            else batch_stats_init.clone().to(p.device)
        d_losses, c_losses = [], []
        for b, (lo, hi) in enumerate(_tree_blocks(g)):
            pb, yb = p[lo:hi], y[lo:hi]
            # label -> the LAST node carrying it (the reference's dict comprehension keeps the last key)
            rows, keys = [], []
            for label in range(1, C):
                idx = (yb == label).nonzero().flatten()
                if idx.numel():
                    node = int(idx[-1])
                    stats[b, label - 1] = pb[node].detach()
                    rows.append(pb[node])
                    keys.append(label - 1)
            cur = torch.stack(rows, 0)
            if self.cached_mean_pos_enc is not None:
                c_loss = ((cur - self.cached_mean_pos_enc[torch.tensor(keys, device=p.device)]) ** 2).sum()
            else:
                c_loss = torch.zeros((), device=p.device)
            aff = torch.exp(-(pb[None, :, :] - pb[:, None, :]).abs().sum(2))
            d_losses.append(F.smooth_l1_loss(aff, torch.exp(-all_pos_encs_cache[b].to(p.device))))
            c_losses.append(c_loss.reshape(()))
        mean_stats = stats.mean(0).detach()
        self.cached_mean_pos_enc = mean_stats if self.cached_mean_pos_enc is None \
            else 0.15 * self.cached_mean_pos_enc + 0.85 * mean_stats
        return torch.stack(d_losses).mean(), torch.stack(c_losses).mean()
