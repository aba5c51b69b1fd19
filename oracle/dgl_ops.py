"""DGL-0.7.x graph container and conv modules restated in plain PyTorch (CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each op is written as the op
sequence DGL issues, over the exact DGL edge list, so it
doubles as the CPU timing baseline.

Reference call sites this follows (the arithmetic itself is DGL's, not in the
reference tree): /root/reference/models.py:8 (import), :172-182 GraphConv,
:301-314 / :425-456 / :506-521 GATConv, :358-383 GINConv, :668-679 SAGEConv;
graph API at job_runner.py:822-838, :1319-1344, :1779-1801, :1390, :1882.

Version switches: ``gat_bias``, ``gat_res_identity_rule``
("in!=F" is 0.7.x, "in!=H*F" is ≥0.8/0.9), ``sage_bias_layout``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kinks

VERSION_SWITCHES = {
    "gat_bias": True,
    "gat_res_identity_rule": "in!=F",
    "sage_bias_layout": "single",
}


# ----------------------------------------------------------------------------
# graph container (DGLGraph subset) — integer work, must be bit-exact
# ----------------------------------------------------------------------------
class Graph:
    """Homogeneous directed multigraph: edge i is ``src[i] -> dst[i]`` (DGL edge id i)."""

    def __init__(self, src, dst, num_nodes, batch_num_nodes=None, batch_num_edges=None):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self.num_nodes = int(num_nodes)
        self.ndata = {}
        self._bnn = torch.tensor([self.num_nodes]) if batch_num_nodes is None else batch_num_nodes
        self._bne = torch.tensor([self.src.numel()]) if batch_num_edges is None else batch_num_edges

    # --- the DGLGraph surface the reference touches ---
    def number_of_nodes(self):
        return self.num_nodes

    def number_of_edges(self):
        return int(self.src.numel())

    def nodes(self):
        return torch.arange(self.num_nodes, dtype=torch.int64)

    def edges(self):
        return self.src, self.dst

    def add_edges(self, u, v):
        u = torch.as_tensor(u, dtype=torch.int64)
        v = torch.as_tensor(v, dtype=torch.int64)
        self.src = torch.cat([self.src, u])
        self.dst = torch.cat([self.dst, v])
        self._bne = torch.tensor([self.src.numel()])

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.num_nodes)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.num_nodes)

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    @property
    def batch_size(self):
        return int(self._bnn.numel())

    def adjacency_dense(self):
        """``g.adjacency_matrix().to_dense()``: A[dst? no — DGL default is A[src, dst] with transpose=False in 0.7]."""
        a = torch.zeros(self.num_nodes, self.num_nodes)
        a[self.src, self.dst] = 1.0
        return a

    def to(self, *_a, **_k):
        return self

    def cpu(self):
        return self


def remove_self_loop(g):
    keep = g.src != g.dst
    out = Graph(g.src[keep], g.dst[keep], g.num_nodes)
    out.ndata = dict(g.ndata)
    return out


def edges_from_adj(adj, symmetric_path="digraph"):
    """Edge list DGL ends up with for ``DGLGraph(nx.DiGraph(adj))`` / ``nx.Graph(adj)``.

    Both construction paths of the reference (job_runner.py:1783-1785 and
    :1336-1340/:834-837) iterate the networkx adjacency dict-of-dicts, which
    ``nx.from_numpy_array`` fills in row-major order of the non-zeros; a
    symmetric ``nx.Graph`` hands DGL the same directed list (both directions,
    source-major).  Verified against the literal networkx calls in
    tests/golden/make_golden.py.
    """
    adj = np.asarray(adj)
    s, d = np.nonzero(adj)
    return s.astype(np.int64), d.astype(np.int64)


def graph_from_adj(adj):
    """job_runner.py:1779-1801 minus the PE call: DiGraph(adj) → strip self loops → append (k,k)."""
    adj = np.asarray(adj)
    n = adj.shape[0]
    s, d = edges_from_adj(adj)
    keep = s != d
    g = Graph(s[keep], d[keep], n)
    g.add_edges(g.nodes(), g.nodes())
    return g


def batch(graphs):
    """``dgl.batch``: disjoint union; node ids shifted by Σ n_j, edge order preserved, ndata concatenated."""
    n_nodes = torch.tensor([g.num_nodes for g in graphs], dtype=torch.int64)
    n_edges = torch.tensor([g.number_of_edges() for g in graphs], dtype=torch.int64)
    off = torch.cumsum(n_nodes, 0) - n_nodes
    src = torch.cat([g.src + o for g, o in zip(graphs, off)])
    dst = torch.cat([g.dst + o for g, o in zip(graphs, off)])
    out = Graph(src, dst, int(n_nodes.sum()), n_nodes, n_edges)
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    return out


def unbatch(g):
    outs = []
    no = eo = 0
    for n, e in zip(g.batch_num_nodes().tolist(), g.batch_num_edges().tolist()):
        h = Graph(g.src[eo:eo + e] - no, g.dst[eo:eo + e] - no, n)
        for k, v in g.ndata.items():
            h.ndata[k] = v[no:no + n]
        outs.append(h)
        no += n
        eo += e
    return outs


# ----------------------------------------------------------------------------
# message-passing primitives (DGL SpMM / SDDMM restated)
# ----------------------------------------------------------------------------
def _seg_sum(e_val, dst, n):
    out = torch.zeros((n,) + e_val.shape[1:], dtype=e_val.dtype)
    return out.index_add_(0, dst, e_val)


def _seg_max(e_val, dst, n):
    out = torch.full((n,) + e_val.shape[1:], -math.inf, dtype=e_val.dtype)
    idx = dst.view(-1, *([1] * (e_val.dim() - 1))).expand_as(e_val)
    return out.scatter_reduce(0, idx, e_val, reduce="amax", include_self=True)


def edge_softmax(g, e):
    """DGL edge_softmax by destination: SpMM(copy_e,max), sub, exp, SpMM(copy_e,sum), div."""
    m = _seg_max(e, g.dst, g.num_nodes)
    ex = torch.exp(e - m[g.dst])
    s = _seg_sum(ex, g.dst, g.num_nodes)
    return ex / s[g.dst]


# ----------------------------------------------------------------------------
# conv modules, DGL parameter names
# ----------------------------------------------------------------------------
class GATConv(nn.Module):
    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0,
                 negative_slope=0.2, residual=False, activation=None,
                 allow_zero_in_degree=False, bias=None):
        super().__init__()
        self._in, self._out, self._heads = in_feats, out_feats, num_heads
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.negative_slope = negative_slope
        self.allow_zero_in_degree = allow_zero_in_degree
        if residual:
            rule = VERSION_SWITCHES["gat_res_identity_rule"]
            needs_linear = (in_feats != out_feats) if rule == "in!=F" else (in_feats != out_feats * num_heads)
            self.res_fc = nn.Linear(in_feats, num_heads * out_feats, bias=False) if needs_linear else nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        use_bias = VERSION_SWITCHES["gat_bias"] if bias is None else bias
        if use_bias:
            self.bias = nn.Parameter(torch.empty(num_heads * out_feats))
        else:
            self.register_buffer("bias", None)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, g, feat, get_attention=False):
        if not self.allow_zero_in_degree and bool((g.in_degrees() == 0).any()):
            raise RuntimeError("There are 0-in-degree nodes in the graph (DGLError in the reference stack)")
        n, H, Fo = g.num_nodes, self._heads, self._out
        h = kinks.dropout(self.feat_drop, feat)                    # K0 (kinks.*: the plain torch op unless a test
        z = self.fc(h).view(n, H, Fo)                              # K1  replays the device's decisions, oracle/kinks.py)
        el = (z * self.attn_l).sum(-1, keepdim=True)               # K2
        er = (z * self.attn_r).sum(-1, keepdim=True)
        e = kinks.leaky_relu(el[g.src] + er[g.dst], self.negative_slope)   # K3, K4
        a = kinks.dropout(self.attn_drop, edge_softmax(g, e))      # K5, K6
        rst = _seg_sum(z[g.src] * a, g.dst, n)                     # K7
        if self.res_fc is not None:                                # K8
            rst = rst + self.res_fc(h).view(n, -1, Fo)
        if self.bias is not None:                                  # K9
            rst = rst + self.bias.view(1, H, Fo)
        if self.activation is not None:
            rst = self.activation(rst)
        return (rst, a) if get_attention else rst


class GraphConv(nn.Module):
    """norm='both', weight [in,out] (xavier_uniform), bias zeros."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        assert norm == "both" and weight and bias
        self._in, self._out = in_feats, out_feats
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.empty(out_feats))
        self._activation = activation
        self.allow_zero_in_degree = allow_zero_in_degree
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, g, feat):
        if not self.allow_zero_in_degree and bool((g.in_degrees() == 0).any()):
            raise RuntimeError("There are 0-in-degree nodes in the graph")
        n = g.num_nodes
        norm = g.out_degrees().to(feat.dtype).clamp(min=1).pow(-0.5)
        x = feat * norm.view(-1, 1)
        if self._in > self._out:
            x = x @ self.weight
            rst = _seg_sum(x[g.src], g.dst, n)
        else:
            rst = _seg_sum(x[g.src], g.dst, n)
            rst = rst @ self.weight
        norm = g.in_degrees().to(feat.dtype).clamp(min=1).pow(-0.5)
        rst = rst * norm.view(-1, 1)
        rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class Block:
    """Message-flow graph of DGL's neighbour sampling (``dgl.to_block``): edges ``src[i] -> dst[i]`` from ``num_src``
    source rows to ``num_dst`` destination rows; destination node i is source node i."""

    def __init__(self, src, dst, num_src, num_dst):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self.num_src, self.num_dst = int(num_src), int(num_dst)
        self.srcdata, self.dstdata = {}, {}


class SAGEConv(nn.Module):
    """aggregator_type='pool' only (models.py:660, 668-679)."""

    def __init__(self, in_feats, out_feats, aggregator_type="pool", feat_drop=0.0, bias=True,
                 norm=None, activation=None):
        super().__init__()
        assert aggregator_type == "pool"
        self._in, self._out = in_feats, out_feats
        self.feat_drop = nn.Dropout(feat_drop)
        self.norm, self.activation = norm, activation
        self.fc_pool = nn.Linear(in_feats, in_feats)
        per_linear = VERSION_SWITCHES["sage_bias_layout"] == "per_linear"
        self.fc_self = nn.Linear(in_feats, out_feats, bias=per_linear)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=per_linear)
        if not per_linear and bias:
            self.bias = nn.Parameter(torch.zeros(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, g, feat):
        h = kinks.dropout(self.feat_drop, feat)
        m = kinks.relu(self.fc_pool(h))
        if isinstance(g, Block):                   # DGL: feat_dst = feat_src[:number_of_dst_nodes()]
            n = g.num_dst
            h = h[:n]
        else:
            n = g.num_nodes
        neigh = kinks.seg_max_nodes(m, g.src, g.dst, n)            # max over in-neighbours; DGL zero-fills empty rows
        rst = self.fc_self(h) + self.fc_neigh(neigh)
        if self.bias is not None:
            rst = rst + self.bias
        if self.activation is not None:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst


class GINConv(nn.Module):
    """aggregator 'mean' (models.py:358-383); eps is a learnable [1] parameter, init 0."""

    def __init__(self, apply_func, aggregator_type, init_eps=0, learn_eps=False):
        super().__init__()
        assert aggregator_type == "mean"
        self.apply_func = apply_func
        if learn_eps:
            self.eps = nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, g, feat):
        n = g.num_nodes
        s = _seg_sum(feat[g.src], g.dst, n)
        deg = g.in_degrees().to(feat.dtype).clamp(min=1)
        neigh = s / deg.view(-1, 1)
        rst = (1 + self.eps) * feat + neigh
        if self.apply_func is not None:
            rst = self.apply_func(rst)
        return rst
