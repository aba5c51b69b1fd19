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
