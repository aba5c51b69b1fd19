"""Wiring of the reference's GNN stacks over the restated DGL convs (CPU oracle).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
/root/reference/models.py:160-194 (GCN), :283-340 (GAT), :343-400 (GIN),
:403-484 (GATPSPGNN), :487-540 (GATPSPGNNNL), :650-696 (SAGE) and the GNN half
of the ``*Net`` wrappers (:196-281, :725-822, :824-933, :936-1047, :1050-1174).
Sub-module names equal the reference's so a state dict moves between the
reference, this oracle and the CUDA product unchanged.  Checked against the
reference's own classes by tests/golden/wiring_*.npz.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kinks
from .dgl_ops import GATConv, GINConv, GraphConv, SAGEConv


class _Stack(nn.Module):
    def reset_parameters(self):
        for m in self.modules():
            if m is not self and isinstance(m, (GATConv, GraphConv, SAGEConv)):
                m.reset_parameters()


class GCN(_Stack):
    # models.py:160-194
    def __init__(self, num_layers, in_dim, num_hiddens, num_classes, activation):
        super().__init__()
        self.num_layers = num_layers
        dims = [in_dim] + list(num_hiddens[:num_layers])
        self.gcn_layers = nn.ModuleList(
            [GraphConv(dims[i], dims[i + 1], activation=activation) for i in range(num_layers)]
            + [GraphConv(dims[-1], num_classes)])

    def forward(self, g):
        h = g.ndata["fvs"]
        for layer in self.gcn_layers:
            h = layer(g, h)
        return h


class GAT(_Stack):
    # models.py:283-340; layer 0 and the output layer have zero drop rates (:301-314)
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, heads, activation, feat_drop,
                 attn_drop, negative_slope, residual, norm=False):
        super().__init__()
        self.num_layers, self.norm = num_layers, norm
        widths = [in_dim] + [num_hiddens[i] * heads[i] for i in range(num_layers)]
        layers = []
        for i in range(num_layers):
            fd, ad = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)
            layers.append(GATConv(widths[i], num_hiddens[i], heads[i], fd, ad, negative_slope, residual, activation))
        layers.append(GATConv(widths[-1], out_ch, heads[num_layers], 0.0, 0.0, negative_slope, residual, None))
        self.gat_layers = nn.ModuleList(layers)

    def forward(self, g):
        h = g.ndata["fvs"]
        for layer in self.gat_layers[:-1]:
            h = layer(g, h).flatten(1)
        h = self.gat_layers[-1](g, h).mean(1)
        return F.normalize(h, p=2, dim=1) if self.norm else h


def _gin_mlp(i, o):
    # kinks.*: nn.Dropout / nn.LeakyReLU unless a test replays the device's decisions (oracle/kinks.py)
    return nn.Sequential(nn.Linear(i, o), kinks.Dropout(0.1), kinks.LeakyReLU(), nn.Linear(o, o), kinks.LeakyReLU())


class GIN(nn.Module):
    # models.py:343-400
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, norm=False):
        super().__init__()
        self.norm = norm
        dims = [in_dim] + list(num_hiddens[:num_layers]) + [out_ch]
        self.gin_layers = nn.ModuleList(
            [GINConv(_gin_mlp(dims[i], dims[i + 1]), "mean", learn_eps=True) for i in range(num_layers + 1)])

    def forward(self, g):
        h = g.ndata["fvs"]
        for layer in self.gin_layers:
            h = layer(g, h)
        return F.normalize(h, p=2, dim=1) if self.norm else h


class SAGE(_Stack):
    # models.py:650-696; layer 0 feat_drop 0, output layer default args
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, node_ks, node_sample_rate=0.3,
                 activation=F.elu, feat_drop=0.1, aggregator_type="pool", norm=None):
        super().__init__()
        dims = [in_dim] + list(num_hiddens[:num_layers])
        layers = [SAGEConv(dims[i], dims[i + 1], aggregator_type=aggregator_type,
                           feat_drop=0.0 if i == 0 else feat_drop, activation=activation, norm=norm)
                  for i in range(num_layers)]
        layers.append(SAGEConv(dims[-1], out_ch, aggregator_type=aggregator_type))
        self.g_layers = nn.ModuleList(layers)

    def forward(self, g):
        h = g.ndata["fvs"]
        for layer in self.g_layers:
            h = layer(g, h)
        return h

    def forward_batch(self, blocks, x):
        # models.py:685-689
        h = x
        for layer, block in zip(self.g_layers, blocks):
            h = layer(block, h)
        return h


class GATPSPGNN(_Stack):
    # models.py:403-484: structure stream consumes h_p BEFORE the position layer updates it (:476-479);
    # output layer keeps the activation (:436-440); pgnn drops are zero on the first and last layer (:443-456).
    def __init__(self, num_layers, in_dim, pos_in_dim, num_hiddens, pos_hiddens, pos_heads, out_ch, heads,
                 activation, feat_drop, attn_drop, negative_slope, residual, norm=False, p_activation=torch.tanh):
        super().__init__()
        L = self.num_layers = num_layers
        s_w = [in_dim] + [num_hiddens[i] * heads[i] for i in range(L)]
        p_w = [pos_in_dim] + [pos_hiddens[i] * pos_heads[i] for i in range(L)]
        gat, pg = [], []
        for i in range(L):
            fd, ad = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)
            gat.append(GATConv(s_w[i] + p_w[i], num_hiddens[i], heads[i], fd, ad, negative_slope, residual, activation))
            fd, ad = (0.0, 0.0) if i in (0, L - 1) else (feat_drop, attn_drop)
            pg.append(GATConv(p_w[i], pos_hiddens[i], pos_heads[i], fd, ad, negative_slope, True, p_activation))
        gat.append(GATConv(s_w[L] + p_w[L], out_ch, heads[L], 0.0, 0.0, negative_slope, residual, activation))
        self.gat_layers, self.pgnn_layers = nn.ModuleList(gat), nn.ModuleList(pg)

    def forward(self, g):
        h_p, h_s = g.ndata["pos_enc"], g.ndata["fvs"]
        for i in range(self.num_layers):
            h_s = self.gat_layers[i](g, torch.cat([h_s, h_p], 1)).flatten(1)
            h_p = self.pgnn_layers[i](g, h_p).flatten(1)
        h_s = self.gat_layers[-1](g, torch.cat([h_s, h_p], 1)).mean(1)
        return h_s, h_p


class GATPSPGNNNL(_Stack):
    # models.py:487-540: no position stream; the initial pos_enc is re-concatenated at every layer
    def __init__(self, num_layers, in_dim, pos_in_dim, num_hiddens, out_ch, heads, activation, feat_drop,
                 attn_drop, negative_slope, residual, norm=False):
        super().__init__()
        L = self.num_layers = num_layers
        s_w = [in_dim] + [num_hiddens[i] * heads[i] for i in range(L)]
        gat = []
        for i in range(L):
            fd, ad = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)
            gat.append(GATConv(s_w[i] + pos_in_dim, num_hiddens[i], heads[i], fd, ad, negative_slope, residual, activation))
        gat.append(GATConv(s_w[L] + pos_in_dim, out_ch, heads[L], 0.0, 0.0, negative_slope, residual, activation))
        self.gat_layers = nn.ModuleList(gat)

    def forward(self, g):
        h_p, h_s = g.ndata["pos_enc"], g.ndata["fvs"]
        for layer in self.gat_layers[:-1]:
            h_s = layer(g, torch.cat([h_s, h_p], 1)).flatten(1)
        h_s = self.gat_layers[-1](g, torch.cat([h_s, h_p], 1)).mean(1)
        return h_s, h_p


class GNNNet(nn.Module):
    """GNN half + ``gnn_out`` head of ``GCNNet/SAGENet/GATNet/GINNet/GATPositionSPGNNNet`` (CNN trunk omitted).

    ``kind`` ∈ {"gcn","sage","gat","gin","spgnn"}; ``model`` is ``settings.MODEL`` (extra keys ignored).
    """

    def __init__(self, kind, model):
        super().__init__()
        m = dict(model)
        self.kind = kind
        fv, emb, out_ch, hid = m["fv_dim"], m["node_embed_dim"], m["out_ch"], m["num_hiddens"]
        if kind == "gcn":
            self.gcn = GCN(m["num_gcn_layers"], fv, hid, emb, F.elu)
        elif kind == "sage":
            self.sage = SAGE(m["num_layers"], fv, hid, emb, node_ks=m["node_ks"],
                             node_sample_rate=m["node_sample_rate"], activation=F.elu, feat_drop=m["feat_drop"],
                             aggregator_type=m.get("aggregator_type", "pool"))
        elif kind == "gin":
            self.gin = GIN(m["num_gin_layers"], fv, hid, emb)
            self.gnn_lobe_out = nn.Linear(emb, 6)      # models.py:988-989 (unused heads, but in the state dict)
            self.gnn_lung_out = nn.Linear(emb, 3)
        elif kind == "gat":
            L = m["num_gat_layers"]
            heads = [m["num_heads"]] * L + [m["num_out_heads"]]
            self.gat = GAT(L, fv, hid, emb, heads, F.elu, m["feat_drop"], m["attn_drop"], m["negative_slope"],
                           m.get("res", True))
        elif kind == "spgnn":
            L = m["num_gat_layers"]
            heads = [m["num_heads"]] * L + [m["num_out_heads"]]
            p_act = torch.tanh if m.get("p_act", "tahn") == "tahn" else F.elu
            if m.get("mode", "PEL") == "PEL":
                self.gat = GATPSPGNN(L, fv, m["pos_enc_dim"], hid, m["pos_hiddens"], [m["num_pos_heads"]] * (L + 1),
                                     emb, heads, F.elu, m["feat_drop"], m["attn_drop"], m["negative_slope"],
                                     m.get("res", True), p_activation=p_act)
            else:
                self.gat = GATPSPGNNNL(L, fv, m["pos_enc_dim"], hid, emb, heads, F.elu, m["feat_drop"],
                                       m["attn_drop"], m["negative_slope"], m.get("res", True))
        else:
            raise ValueError(kind)
        self.gnn_out = nn.Linear(emb, out_ch)

    @property
    def stack(self):
        return getattr(self, {"gcn": "gcn", "sage": "sage", "gin": "gin", "gat": "gat", "spgnn": "gat"}[self.kind])

    def init_like_reference(self):
        """models.py:896-900 / :1142-1146 + initializer.py:17-30 (GIN keeps torch's Linear default, :1008-1011)."""
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                mod.reset_parameters()
        if self.kind != "gin":
            self.stack.reset_parameters()
        nn.init.xavier_normal_(self.gnn_out.weight, gain=nn.init.calculate_gain("linear"))
        nn.init.constant_(self.gnn_out.bias, 0.0)

    def forward_batch(self, blocks, x):
        # SAGENet.forward_batch, models.py:814-817
        n_embed = self.stack.forward_batch(blocks, x)
        return self.gnn_out(n_embed), n_embed

    def forward(self, g):
        res = self.stack(g)
        if self.kind == "spgnn":
            return self.gnn_out(res[0]), res[0], res[1]
        return self.gnn_out(res), res


def cross_entropy_masked(logits, y, mask, class_weights):
    """job_runner.py:1900 — ``F.cross_entropy(out[mask], y[mask], weight)`` = Σ wᵢℓᵢ / Σ wᵢ."""
    return F.cross_entropy(logits[mask], y[mask], weight=class_weights)


def decide_per_tree(logits, batch_num_nodes):
    """job_runner.py:158-165 — per class 1..21, the node with the highest softmax probability (first max)."""
    out = []
    o = 0
    for n in batch_num_nodes.tolist():
        p = F.softmax(logits[o:o + n], dim=1)
        out.append(torch.max(p[:, 1:], 0)[1] + o)
        o += n
    return torch.stack(out)
