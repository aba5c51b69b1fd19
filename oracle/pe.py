"""Positional-encoding inits restated on CPU (numpy) — TEST INFRASTRUCTURE.

Follows /root/reference/job_runner.py:
  :1727-1757 get_anchors_from_cnn_prediction   :1712-1725 add_distal_leafs
  :1759-1777 generate_distant_pos_enc (live)    :1684-1702 generate_rw_pos_enc (dormant)

Tie rule.  The reference's distal-leaf choice among equally far leaves depends
on CPython ``set`` iteration order (``nx.descendants`` returns a set, then a
stable sort takes the last element, :1718-1724) and is not reproducible as
written.  ``tie_rule="max_index"`` (deterministic: largest node index among the
farthest leaves) is what the device kernel implements; ``tie_rule="reference"``
runs the literal networkx calls so a test can bound the difference.
"""
from __future__ import annotations

from collections import deque

import numpy as np


def softmax_rows(x):
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def select_anchors(fvs_out, n_labels=21):
    """:1733-1741 — for label 1..21: argmax over nodes of P[:,label]*mask (first max), then mask it out."""
    p = softmax_rows(fvs_out)
    mask = np.ones(p.shape[0], dtype=np.float64)   # np.ones_like(int64 y) * 1.0 → float64
    anchors = []
    for label in range(1, n_labels + 1):
        idx = int(np.argmax(p[:, label] * mask))
        mask[idx] = 0.0
        anchors.append(idx)
    return anchors


def _children_lists(adj):
    up = np.triu(np.asarray(adj), k=1)
    return [np.nonzero(up[i])[0].tolist() for i in range(up.shape[0])]


def distal_leaves(anchors, adj, tie_rule="max_index"):
    """:1712-1725 — per anchor the farthest descendant leaf in the DAG triu(adj); the anchor itself if it has none."""
    if tie_rule == "reference":
        import networkx as nx
        G = nx.DiGraph(np.triu(np.asarray(adj)))
        G.remove_edges_from(nx.selfloop_edges(G))
        out = []
        for a in anchors:
            leafs = {n: nx.shortest_path_length(G, a, n) for n in nx.descendants(G, a) if G.out_degree(n) == 0}
            out.append(a if len(leafs) == 0 else sorted(leafs.items(), key=lambda x: x[1])[-1][0])
        return out
    ch = _children_lists(adj)
    out = []
    for a in anchors:
        dist = {a: 0}
        q = deque([a])
        while q:
            u = q.popleft()
            for v in ch[u]:
                if v not in dist:
                    dist[v] = dist[u] + 1
                    q.append(v)
        best, best_d = a, -1
        for v, d in dist.items():
            if v != a and len(ch[v]) == 0 and (d > best_d or (d == best_d and v > best)):
                best, best_d = v, d
        out.append(best)
    return out


def anchors_39(fvs_out, adj, pos_enc_dim=39, tie_rule="max_index"):
    """:1727-1757 — 21 CNN anchors (+ distal leaves of the first 18 when POS_ENC_DIM == 39)."""
    anchors = select_anchors(fvs_out)
    if pos_enc_dim == 39:
        return anchors + distal_leaves(anchors[:-3], adj, tie_rule)
    if pos_enc_dim == 21:
        return anchors
    raise NotImplementedError(f"pos enc dim : {pos_enc_dim}!")


def _bfs_all(nbrs, s):
    n = len(nbrs)
    d = np.full(n, -1, dtype=np.int64)
    d[s] = 0
    q = deque([s])
    while q:
        u = q.popleft()
        for v in nbrs[u]:
            if d[v] < 0:
                d[v] = d[u] + 1
                q.append(v)
    return d


def hop_matrix(adj):
    """All-pairs hop counts on the self-loop-free graph (nx.all_pairs_shortest_path_length, :1764)."""
    a = np.asarray(adj).copy()
    np.fill_diagonal(a, 0)
    nbrs = [np.nonzero(a[i])[0].tolist() for i in range(a.shape[0])]
    return np.stack([_bfs_all(nbrs, s) for s in range(a.shape[0])])


def dist_pos_enc(adj, anchors):
    """:1759-1777 — pos_enc[n,k] = hops(n, anchor_k) / diameter, float32; also the [n,n] all-pairs matrix."""
    hops = hop_matrix(adj)
    if (hops < 0).any():
        raise ValueError("graph is not connected (nx.diameter raises in the reference)")
    diameter = int(hops.max())
    # python float division then float32 store == fp32(int/int in fp64); int/int ≤ 2^24 ⇒ same as fp32 divide
    pe = (hops[:, anchors].astype(np.float64) / float(diameter)).astype(np.float32)
    all_pe = (hops.astype(np.float64) / float(diameter)).astype(np.float32)
    return pe, all_pe, diameter


def rw_pos_enc(adj, pos_enc_dim=39):
    """:1684-1702 — diag((A D^-1)^k), k = 1..pos_enc_dim, float64 → float32; A without self loops."""
    a = np.asarray(adj).astype(np.float64)
    np.fill_diagonal(a, 0.0)
    indeg = a.sum(axis=0)
    dinv = np.eye(a.shape[0]) * np.clip(indeg, 1, None) ** -1.0
    m = a @ dinv
    cols = [np.diagonal(m).astype(np.float32)]
    mp = m
    for _ in range(pos_enc_dim - 1):
        mp = mp @ m
        cols.append(np.diagonal(mp).astype(np.float32))
    return np.stack(cols, axis=-1)


# ----------------------------------------------------------------------------
# dormant extras, literal restatements on the oracle Graph
# ----------------------------------------------------------------------------
def _norm_laplacian(g):
    """job_runner.py:1632-1634 / :1814-1818: L = I - N A N with A[src, dst] and N = diag(clip(in_deg, 1)^-1/2)."""
    n = g.num_nodes
    a = np.zeros((n, n))
    a[g.src.numpy(), g.dst.numpy()] = 1.0
    d = np.clip(g.in_degrees().numpy(), 1, None) ** -0.5
    return np.eye(n) - d[:, None] * a * d[None, :]


def eigen_basis(g, pos_enc_dim=39):
    """job_runner.py:1630-1645 compute_eigen_basis for ONE graph: (eigenvalues ascending, eigvec[:, 1:dim+1] fp32,
    zero-padded to pos_enc_dim columns when n <= pos_enc_dim)."""
    val, vec = np.linalg.eig(_norm_laplacian(g))
    idx = val.argsort()
    val, vec = np.real(val[idx]), np.real(vec[:, idx])
    out = vec[:, 1:pos_enc_dim + 1].astype(np.float32)
    if out.shape[1] < pos_enc_dim:
        out = np.pad(out, ((0, 0), (0, pos_enc_dim - out.shape[1])))
    return val, out


def laplacian_pos_loss(graphs, ps, lamb, pos_enc_dim):
    """job_runner.py:1803-1825 over a list of (unbatched) oracle graphs and their [n, k] position embeddings."""
    import torch
    total = []
    for g, p in zip(graphs, ps):
        pz = p - torch.mean(p, dim=0, keepdim=True).detach()
        pn = pz / (torch.std(p, dim=0, keepdim=True) + 1e-7).detach()
        n = g.num_nodes
        L = torch.tensor(_norm_laplacian(g), dtype=p.dtype)
        pT = torch.transpose(pn, 1, 0)
        loss1 = torch.trace(torch.mm(torch.mm(pT, L), pn))
        ptp = torch.mm(pT, pn) - torch.eye(pn.shape[1], dtype=p.dtype)
        total.append((loss1 + lamb * torch.norm(ptp, p="fro")) / (pos_enc_dim * n))
    return torch.stack(total).mean()


class DistPosLoss:
    """job_runner.py:1827-1861 dist_pos_loss with its ``cached_mean_pos_enc`` state; ``batch_stats`` replaces the
    reference's ``torch.rand`` initial fill so that two implementations can be compared."""

    def __init__(self, nr_class=22, pos_enc_dim=39):
        self.nr_class, self.pos_enc_dim, self.cached_mean_pos_enc = nr_class, pos_enc_dim, None

    def __call__(self, ps, ys, all_pos_encs_cache, batch_stats):
        import torch
        import torch.nn.functional as F
        batch_stats = batch_stats.clone()
        total_d, total_c = [], []
        for b, (p, y) in enumerate(zip(ps, ys)):
            label_mapping = {y[n].item(): n for n in range(y.shape[0]) if y[n].item() != 0}
            existing = sorted(set(range(1, self.nr_class)) & set(label_mapping.keys()))
            cur = []
            for label in range(1, self.nr_class):
                if label in label_mapping:
                    batch_stats[b, label - 1, ::] = p[label_mapping[label]].detach()
                    cur.append(p[label_mapping[label]])
            cur = torch.stack(cur, dim=0)
            if self.cached_mean_pos_enc is not None:
                c_loss = ((cur - self.cached_mean_pos_enc[(np.asarray(existing) - 1)]) ** 2).sum()
            else:
                c_loss = torch.zeros(())
            n = p.shape[0]
            x = p.unsqueeze(0).repeat(n, 1, 1)
            yy = p.unsqueeze(1).repeat(1, n, 1)
            aff = torch.exp(-1.0 * torch.abs(x - yy).sum(dim=2))
            total_d.append(F.smooth_l1_loss(aff, torch.exp(-all_pos_encs_cache[b])))
            total_c.append(c_loss.reshape(()))
        mean_stats = batch_stats.mean(dim=0).detach()
        self.cached_mean_pos_enc = mean_stats if self.cached_mean_pos_enc is None \
            else 0.15 * self.cached_mean_pos_enc + 0.85 * mean_stats
        return torch.stack(total_d).mean(), torch.stack(total_c).mean()
