"""Independent dense-adjacency formulations used to cross-check oracle.dgl_ops.

TEST INFRASTRUCTURE.  Nothing here shares code with dgl_ops: attention is a
masked softmax over an N×N matrix, GCN is D^-1/2 A D^-1/2 X W, SAGE-pool a masked
max, GIN a row-normalised matmul.  ``adj[u, v] = 1`` means an edge u → v (simple
graphs only).  Parameters are passed explicitly so the check does not depend on
module plumbing.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gat_dense(adj, x, w_fc, attn_l, attn_r, w_res, bias, negative_slope, activation, heads, out_feats):
    n = x.shape[0]
    z = (x @ w_fc.t()).view(n, heads, out_feats)
    el = torch.einsum("nhf,hf->nh", z, attn_l.view(heads, out_feats))
    er = torch.einsum("nhf,hf->nh", z, attn_r.view(heads, out_feats))
    s = el.unsqueeze(1) + er.unsqueeze(0)                      # s[u, v, h]
    s = F.leaky_relu(s, negative_slope)
    s = s.masked_fill(adj.unsqueeze(-1) == 0, float("-inf"))
    a = torch.softmax(s, dim=0)                                # over sources u of each v
    out = torch.einsum("uvh,uhf->vhf", a, z)
    if w_res is not None:                                      # a [HF,D] matrix, or the string "identity"
        res = x if isinstance(w_res, str) else x @ w_res.t()
        out = out + res.view(n, -1, out_feats)
    if bias is not None:
        out = out + bias.view(1, heads, out_feats)
    return activation(out) if activation is not None else out


def gcn_dense(adj, x, weight, bias, activation):
    dout = adj.sum(1).clamp(min=1).pow(-0.5)
    din = adj.sum(0).clamp(min=1).pow(-0.5)
    a_hat = din.view(-1, 1) * adj.t() * dout.view(1, -1)      # [v, u]
    out = a_hat @ x @ weight + bias
    return activation(out) if activation is not None else out


def sage_pool_dense(adj, x, w_pool, b_pool, w_self, w_neigh, bias, activation):
    m = F.relu(x @ w_pool.t() + b_pool)
    big = m.unsqueeze(1).expand(-1, adj.shape[0], -1).masked_fill(adj.unsqueeze(-1) == 0, float("-inf"))
    neigh = big.max(0)[0]
    neigh = torch.where(torch.isinf(neigh), torch.zeros_like(neigh), neigh)
    out = x @ w_self.t() + neigh @ w_neigh.t()
    if bias is not None:
        out = out + bias
    return activation(out) if activation is not None else out


def gin_mean_dense(adj, x, eps):
    deg = adj.sum(0).clamp(min=1)
    return (1 + eps) * x + (adj.t() @ x) / deg.view(-1, 1)
