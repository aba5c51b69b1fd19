"""GNN model classes of the reference, same names / constructor signatures / state-dict keys, on CUDA kernels.

Drop-in for the GNN half of /root/reference/models.py: ``GCN`` (:160-194), ``GAT`` (:283-340), ``GIN``
(:343-400), ``GATPSPGNN`` (:403-484), ``GATPSPGNNNL`` (:487-540), ``SAGE`` (:650-696) and the wrappers
``GCNNet`` (:196-281), ``SAGENet`` (:725-822), ``GATNet`` (:824-933), ``GINNet`` (:936-1047),
``GATPositionSPGNNNet`` (:1050-1174).  ``forward(g, h=None, p=None)`` reads ``g.ndata['fvs']`` /
``g.ndata['pos_enc']`` when ``h`` / ``p`` are omitted (the reference signature is ``forward(g)``).

The 3-D CNN trunk of the wrappers (stage 1) is out of scope: the wrappers accept and ignore its constructor
arguments so ``Net(**settings.MODEL)`` works unchanged, and ``forward_without_gnn`` raises.

Differences from a line-by-line port: ``torch.cat([h_s, h_p])`` never materialises (two-source projection),
``.flatten(1)`` is a no-op (layers produce [N, H·F] directly) and the output layer's ``.mean(1)`` is fused
into the aggregation kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn as snn, ops, stack
from ._lib import SpgnnError

# True: GAT-family stacks run through the planes pipeline (stack.py: TMA-fed tcgen05 projections over split-bf16
# activations, layer kernels that write the next projection's operands).  False: one kernel family per op (ops.py).
USE_STACK = True


def set_trainable(model, trainable):
    for _, parameter in model.named_parameters():
        parameter.requires_grad = trainable


def _feats(g, h, key="fvs"):
    """Layer-0 input: fp32 with 16-byte aligned rows (what the tensor-core projection's 128-bit loads need); a
    tensor with an odd row stride (e.g. a contiguous [N,39] pos_enc) is re-laid out once and cached on the graph."""
    x = g.ndata[key] if h is None else h
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() == 2 and (x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0):
        cache = g.__dict__.setdefault("_aligned_feats", {})
        hit = cache.get(key)
        if hit is None or hit[0] is not x:
            y = ops.empty_padded(x.shape[0], x.shape[1], x.device)
            y.copy_(x)
            cache[key] = hit = (x, y)
        x = hit[1]
    return x


class _ConvStack(nn.Module):
    _resettable = (snn.GATConv, snn.GraphConv, snn.SAGEConv)

    def _stack_plan(self):
        """StackPlan of this module's GATConv wiring, or None when the planes pipeline does not cover it."""
        plan = self.__dict__.get("_plan_cache")
        if plan is None:
            plan = self._build_plan()
            self.__dict__["_plan_cache"] = plan if plan.supported() else False
            plan = self.__dict__["_plan_cache"]
        return plan or None

    def _run_stack(self, g, ext, head):
        """(logits or None, outputs) through the planes pipeline, or None when it does not apply."""
        if not USE_STACK or any(t.requires_grad or not t.is_cuda for t in ext.values()):
            return None
        plan = self._stack_plan()
        if plan is None:
            return None
        if any(not L.conv._allow_zero_in_degree for L in plan.layers):
            g.check_no_zero_in_degree()
        return stack.run_stack(plan, g, ext, self.training, head)

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, self._resettable):
                m.reset_parameters()


class GCN(_ConvStack):
    def __init__(self, num_layers, in_dim, num_hiddens, num_classes, activation):
        super().__init__()
        self.num_layers = num_layers
        dims = [in_dim] + [num_hiddens[i] for i in range(num_layers)]
        self.gcn_layers = nn.ModuleList(
            [snn.GraphConv(a, b, activation=activation) for a, b in zip(dims[:-1], dims[1:])])
        self.gcn_layers.append(snn.GraphConv(dims[-1], num_classes))

    def forward(self, g, h=None):
        h = _feats(g, h)
        for conv in self.gcn_layers:
            h = conv(g, h)
        return h


class GAT(_ConvStack):
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, heads, activation, feat_drop, attn_drop,
                 negative_slope, residual, norm=False):
        super().__init__()
        self.num_layers, self.activation, self.out_ch, self.norm = num_layers, activation, out_ch, norm
        self.gat_layers = nn.ModuleList()
        width = in_dim
        for i in range(num_layers):
            drops = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)      # first layer never drops
            self.gat_layers.append(snn.GATConv(width, num_hiddens[i], heads[i], *drops, negative_slope, residual,
                                               activation))
            width = num_hiddens[i] * heads[i]
        self.gat_layers.append(snn.GATConv(width, out_ch, heads[num_layers], 0.0, 0.0, negative_slope, residual, None))

    def _build_plan(self):
        n = len(self.gat_layers)
        names = ["fvs"] + [f"h{i + 1}" for i in range(n - 1)] + ["emb"]
        layers = [stack.LayerPlan(conv, [names[i]], names[i + 1], mean_heads=(i == n - 1))
                  for i, conv in enumerate(self.gat_layers)]
        return stack.StackPlan(layers, {"fvs": self.gat_layers[0]._in_feats}, ["emb"])

    def forward(self, g, h=None, _head=None):
        h = _feats(g, h)
        res = None if self.norm else self._run_stack(g, {"fvs": h}, _head)
        if res is not None:
            return (res[0], res[1][0]) if _head is not None else res[1][0]
        for conv in self.gat_layers[:-1]:
            h = conv.forward_flat(g, h)
        h = self.gat_layers[-1].forward_flat(g, h, mean_heads=True)
        h = F.normalize(h, p=2, dim=1) if self.norm else h
        return (_head(h), h) if _head is not None else h


class GIN(nn.Module):
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, norm=False):
        super().__init__()
        self.num_layers, self.in_dim, self.out_ch, self.norm = num_layers, in_dim, out_ch, norm
        dims = [in_dim] + [num_hiddens[i] for i in range(num_layers)] + [out_ch]
        self.gin_layers = nn.ModuleList()
        for a, b in zip(dims[:-1], dims[1:]):
            # models.py:358-383: Linear -> Dropout(0.1) -> LeakyReLU -> Linear -> LeakyReLU.  Both activations run in
            # the projections' epilogues; LeakyReLU commutes with dropout (a non-negative per-element scale), so
            # leaky(drop(lin(x))) == drop(leaky(lin(x))).  The placeholders keep the Sequential indices 0 and 3.
            mlp = nn.Sequential(snn.Linear(a, b, act="leaky_relu", slope=0.01), snn.Dropout(0.1),
                                snn.LeakyReLU(fused=True), snn.Linear(b, b, act="leaky_relu", slope=0.01),
                                snn.LeakyReLU(fused=True))
            self.gin_layers.append(snn.GINConv(mlp, "mean", learn_eps=True))

    def forward(self, g, h=None):
        h = _feats(g, h)
        for conv in self.gin_layers:
            h = conv(g, h)
        return F.normalize(h, p=2, dim=1) if self.norm else h


class SAGE(_ConvStack):
    def __init__(self, num_layers, in_dim, num_hiddens, out_ch, node_ks, node_sample_rate=0.3, activation=F.elu,
                 feat_drop=0.1, aggregator_type="pool", norm=None):
        super().__init__()
        self.num_layers, self.node_ks, self.node_sample_rate, self.out_ch = num_layers, node_ks, node_sample_rate, out_ch
        self.g_layers = nn.ModuleList()
        width = in_dim
        for i in range(num_layers):
            self.g_layers.append(snn.SAGEConv(width, num_hiddens[i], aggregator_type=aggregator_type,
                                              feat_drop=0.0 if i == 0 else feat_drop, activation=activation, norm=norm))
            width = num_hiddens[i]
        self.g_layers.append(snn.SAGEConv(width, out_ch, aggregator_type=aggregator_type))

    def forward(self, g, h=None):
        h = _feats(g, h)
        for conv in self.g_layers:
            h = conv(g, h)
        return h

    def forward_batch(self, blocks, x):
        """models.py:685-689 on the blocks of ``sampling.MultiLayerNeighborSampler`` (one block per layer)."""
        if len(blocks) != len(self.g_layers):
            raise SpgnnError(f"SAGE.forward_batch: {len(self.g_layers)} layers need as many blocks, got {len(blocks)}")
        h = x
        for conv, block in zip(self.g_layers, blocks):
            h = conv(block, h)
        return h


class GATPSPGNN(_ConvStack):
    """Two-stream SPGNN: a structure GAT stack fed [h_s | h_p] and a position GAT stack fed h_p."""

    def __init__(self, num_layers, in_dim, pos_in_dim, num_hiddens, pos_hiddens, pos_heads, out_ch, heads, activation,
                 feat_drop, attn_drop, negative_slope, residual, norm=False, p_activation=torch.tanh):
        super().__init__()
        self.num_layers, self.activation, self.pos_hiddens, self.out_ch, self.norm = \
            num_layers, activation, pos_hiddens, out_ch, norm
        self.gat_layers, self.pgnn_layers = nn.ModuleList(), nn.ModuleList()
        s_w, p_w = in_dim, pos_in_dim
        for i in range(num_layers):
            drops = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)
            self.gat_layers.append(snn.GATConv(s_w + p_w, num_hiddens[i], heads[i], *drops, negative_slope, residual,
                                               activation))
            p_drops = (0.0, 0.0) if i in (0, num_layers - 1) else (feat_drop, attn_drop)
            self.pgnn_layers.append(snn.GATConv(p_w, pos_hiddens[i], pos_heads[i], *p_drops, negative_slope, True,
                                                p_activation))
            s_w, p_w = num_hiddens[i] * heads[i], pos_hiddens[i] * pos_heads[i]
        # the output layer keeps the activation (unlike GAT)
        self.gat_layers.append(snn.GATConv(s_w + p_w, out_ch, heads[num_layers], 0.0, 0.0, negative_slope, residual,
                                           activation))

    def _build_plan(self):
        n = self.num_layers
        s_names = ["fvs"] + [f"s{i + 1}" for i in range(n)]
        p_names = ["pos_enc"] + [f"p{i + 1}" for i in range(n)]
        layers = []
        for i in range(n):
            layers.append(stack.LayerPlan(self.gat_layers[i], [s_names[i], p_names[i]], s_names[i + 1]))
            layers.append(stack.LayerPlan(self.pgnn_layers[i], [p_names[i]], p_names[i + 1]))
        layers.append(stack.LayerPlan(self.gat_layers[n], [s_names[n], p_names[n]], "emb", mean_heads=True))
        ext = {"fvs": self.gat_layers[0]._in_feats - self.pgnn_layers[0]._in_feats,
               "pos_enc": self.pgnn_layers[0]._in_feats}
        return stack.StackPlan(layers, ext, ["emb", p_names[n]])

    def forward(self, g, h=None, p=None, _head=None):
        h_p, h_s = _feats(g, p, "pos_enc"), _feats(g, h)
        res = self._run_stack(g, {"fvs": h_s, "pos_enc": h_p}, _head)
        if res is not None:
            return (res[0], res[1][0], res[1][1]) if _head is not None else (res[1][0], res[1][1])
        for s_conv, p_conv in zip(self.gat_layers[:-1], self.pgnn_layers):
            h_s = s_conv.forward_flat(g, h_s, h_p)     # consumes h_p BEFORE the position layer updates it
            h_p = p_conv.forward_flat(g, h_p)
        h_s = self.gat_layers[-1].forward_flat(g, h_s, h_p, mean_heads=True)
        return (_head(h_s), h_s, h_p) if _head is not None else (h_s, h_p)


class GATPSPGNNNL(_ConvStack):
    """SPGNN without the position stream: the initial pos_enc is re-attached at every layer."""

    def __init__(self, num_layers, in_dim, pos_in_dim, num_hiddens, out_ch, heads, activation, feat_drop, attn_drop,
                 negative_slope, residual, norm=False):
        super().__init__()
        self.num_layers, self.activation, self.out_ch, self.norm = num_layers, activation, out_ch, norm
        self.gat_layers = nn.ModuleList()
        s_w = in_dim
        for i in range(num_layers):
            drops = (0.0, 0.0) if i == 0 else (feat_drop, attn_drop)
            self.gat_layers.append(snn.GATConv(s_w + pos_in_dim, num_hiddens[i], heads[i], *drops, negative_slope,
                                               residual, activation))
            s_w = num_hiddens[i] * heads[i]
        self.gat_layers.append(snn.GATConv(s_w + pos_in_dim, out_ch, heads[num_layers], 0.0, 0.0, negative_slope,
                                           residual, activation))

    def _build_plan(self):
        n = len(self.gat_layers)
        names = ["fvs"] + [f"s{i + 1}" for i in range(n - 1)] + ["emb"]
        layers = [stack.LayerPlan(conv, [names[i], "pos_enc"], names[i + 1], mean_heads=(i == n - 1))
                  for i, conv in enumerate(self.gat_layers)]
        pw = self.gat_layers[1]._in_feats - self.gat_layers[0]._out_feats * self.gat_layers[0]._num_heads
        return stack.StackPlan(layers, {"fvs": self.gat_layers[0]._in_feats - pw, "pos_enc": pw}, ["emb"])

    def forward(self, g, h=None, p=None, _head=None):
        h_p, h_s = _feats(g, p, "pos_enc"), _feats(g, h)
        res = self._run_stack(g, {"fvs": h_s, "pos_enc": h_p}, _head)
        if res is not None:
            return (res[0], res[1][0], h_p) if _head is not None else (res[1][0], h_p)
        for conv in self.gat_layers[:-1]:
            h_s = conv.forward_flat(g, h_s, h_p)
        h_s = self.gat_layers[-1].forward_flat(g, h_s, h_p, mean_heads=True)
        return (_head(h_s), h_s, h_p) if _head is not None else (h_s, h_p)


# ------------------------------------------------------------------------------------------------
# *Net wrappers: GNN + nn.Linear head.  CNN-trunk constructor arguments are accepted and ignored.
# ------------------------------------------------------------------------------------------------
class _GNNNet(nn.Module):
    _gnn_attr = None

    def _finish(self, node_embed_dim, out_ch):
        self.gnn_out = snn.Linear(node_embed_dim, out_ch)
        self.out_ch = out_ch

    @property
    def _gnn(self):
        return getattr(self, self._gnn_attr)

    def set_gcn_only(self):
        set_trainable(self, False)
        set_trainable(self._gnn, True)
        set_trainable(self.gnn_out, True)

    def set_all(self):
        set_trainable(self, True)

    def set_cnn_only(self):
        raise SpgnnError("the CNN trunk (stage 1) is not part of spgnn_b200")

    def init(self, initializer=None):
        """models.py:896-900 — HeNorm resets every nn.Linear, then the DGL layers re-init, then the head."""
        if initializer is not None:
            initializer.initialize(self)
        else:
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    m.reset_parameters()
        if hasattr(self._gnn, "reset_parameters"):
            self._gnn.reset_parameters()
        nn.init.xavier_normal_(self.gnn_out.weight, gain=nn.init.calculate_gain("linear"))
        nn.init.constant_(self.gnn_out.bias, 0.0)

    def forward_without_gnn(self, x):
        raise SpgnnError("the CNN trunk (stage 1) is not part of spgnn_b200")

    extract_feature = forward_without_gnn

    def forward(self, g, h=None):
        n_embed = self._gnn(g, h)
        return self.gnn_out(n_embed), n_embed

    def forward_emb(self, g, h=None):
        return self._gnn(g, h)


class GCNNet(_GNNNet):
    _gnn_attr = "gcn"

    def __init__(self, n_layers=None, num_gcn_layers=3, in_ch_list=None, base_ch_list=None, end_ch_list=None,
                 checkpoint_layers=None, kernel_sizes=None, out_ch=22, padding_list=None, conv_strides=None,
                 dropout=0.0, spatial_size=None, fv_dim=1024, num_hiddens=(256, 128, 64), node_embed_dim=1024,
                 norm_method="bn", act_method="relu"):
        super().__init__()
        self.fv_dim, self.num_hiddens, self.node_embed_dim = fv_dim, num_hiddens, node_embed_dim
        self.gcn = GCN(num_layers=num_gcn_layers, in_dim=fv_dim, num_hiddens=num_hiddens, num_classes=node_embed_dim,
                       activation=F.elu)
        self._finish(node_embed_dim, out_ch)


class SAGENet(_GNNNet):
    _gnn_attr = "sage"

    def __init__(self, n_layers=None, num_layers=3, in_ch_list=None, base_ch_list=None, end_ch_list=None,
                 checkpoint_layers=None, kernel_sizes=None, out_ch=22, padding_list=None, conv_strides=None,
                 dropout=0.0, feat_drop=0.1, spatial_size=None, fv_dim=1024, num_hiddens=(256, 128, 64),
                 node_embed_dim=1024, node_ks=(2, 2, 2, 2), node_sample_rate=0.3, aggregator_type="pool",
                 norm_method="bn", act_method="relu"):
        super().__init__()
        self.fv_dim, self.num_hiddens, self.node_embed_dim = fv_dim, num_hiddens, node_embed_dim
        self.sage = SAGE(num_layers=num_layers, in_dim=fv_dim, num_hiddens=num_hiddens, out_ch=node_embed_dim,
                         activation=F.elu, feat_drop=feat_drop, node_ks=node_ks, aggregator_type=aggregator_type,
                         node_sample_rate=node_sample_rate)
        self._finish(node_embed_dim, out_ch)

    def forward_batch(self, blocks, x):
        """models.py:814-817: (n_out, n_embed) for the destination nodes of the last block."""
        n_embed = self.sage.forward_batch(blocks, x)
        return self.gnn_out(n_embed), n_embed


class GATNet(_GNNNet):
    _gnn_attr = "gat"

    def __init__(self, n_layers=None, num_gat_layers=3, num_heads=2, num_out_heads=2, in_ch_list=None,
                 base_ch_list=None, end_ch_list=None, checkpoint_layers=None, kernel_sizes=None, out_ch=22,
                 padding_list=None, conv_strides=None, dropout=0.0, feat_drop=0.1, attn_drop=0.1, negative_slope=0.2,
                 spatial_size=None, fv_dim=1024, num_hiddens=(256, 128, 64), node_embed_dim=1024, res=True,
                 norm_method="bn", act_method="relu"):
        super().__init__()
        self.fv_dim, self.num_hiddens, self.node_embed_dim, self.res = fv_dim, num_hiddens, node_embed_dim, res
        heads = [num_heads] * num_gat_layers + [num_out_heads]
        self.gat = GAT(num_layers=num_gat_layers, in_dim=fv_dim, num_hiddens=num_hiddens, out_ch=node_embed_dim,
                       heads=heads, activation=F.elu, feat_drop=feat_drop, attn_drop=attn_drop,
                       negative_slope=negative_slope, residual=res)
        self._finish(node_embed_dim, out_ch)

    def forward(self, g, h=None):
        return self.gat(g, h, _head=self.gnn_out)        # (n_out, n_embed); the head shares the stack's planes


class GINNet(_GNNNet):
    _gnn_attr = "gin"

    def __init__(self, n_layers=None, num_gin_layers=3, in_ch_list=None, base_ch_list=None, end_ch_list=None,
                 checkpoint_layers=None, kernel_sizes=None, out_ch=22, padding_list=None, conv_strides=None,
                 dropout=0.0, spatial_size=None, fv_dim=1024, num_hiddens=(256, 128, 64), node_embed_dim=1024,
                 norm_method="bn", act_method="relu"):
        super().__init__()
        self.fv_dim, self.num_hiddens, self.node_embed_dim = fv_dim, num_hiddens, node_embed_dim
        self.gin = GIN(num_layers=num_gin_layers, in_dim=fv_dim, num_hiddens=num_hiddens, out_ch=node_embed_dim)
        self._finish(node_embed_dim, out_ch)
        self.gnn_lobe_out = snn.Linear(node_embed_dim, 6)     # present in the reference state dict, never used
        self.gnn_lung_out = snn.Linear(node_embed_dim, 3)

    def set_gcn_only(self):
        super().set_gcn_only()
        set_trainable(self.gnn_lobe_out, True)
        set_trainable(self.gnn_lung_out, True)

    def init(self, initializer=None):
        # models.py:1008-1011: no DGL reset for GIN — the MLPs keep torch's default Linear init
        if initializer is not None:
            initializer.initialize(self)
        nn.init.xavier_normal_(self.gnn_out.weight, gain=nn.init.calculate_gain("linear"))
        nn.init.constant_(self.gnn_out.bias, 0.0)


class GATPositionSPGNNNet(_GNNNet):
    _gnn_attr = "gat"

    def __init__(self, n_layers=None, num_gat_layers=3, num_heads=2, num_out_heads=2, in_ch_list=None,
                 base_ch_list=None, end_ch_list=None, checkpoint_layers=None, kernel_sizes=None, out_ch=22,
                 padding_list=None, conv_strides=None, dropout=0.0, feat_drop=0.1, attn_drop=0.1, negative_slope=0.2,
                 spatial_size=None, fv_dim=1024, num_hiddens=(256, 128, 64), pos_hiddens=(256, 128, 64),
                 num_pos_heads=1, node_embed_dim=1024, pos_enc_dim=39, encodng_merge="cat", norm=False, res=True,
                 norm_method="bn", act_method="relu", p_act="tahn", mode="PEL"):
        super().__init__()
        self.fv_dim, self.pos_enc_dim, self.num_pos_heads, self.mode, self.res = \
            fv_dim, pos_enc_dim, num_pos_heads, mode, res
        self.num_hiddens, self.pos_hiddens, self.node_embed_dim = num_hiddens, pos_hiddens, node_embed_dim
        self.p_act = torch.tanh if p_act == "tahn" else F.elu          # sic: the reference's default spelling
        heads = [num_heads] * num_gat_layers + [num_out_heads]
        pos_heads = [num_pos_heads] * (num_gat_layers + 1)
        if mode == "PEL":
            self.gat = GATPSPGNN(num_layers=num_gat_layers, in_dim=fv_dim, pos_in_dim=pos_enc_dim,
                                 num_hiddens=num_hiddens, pos_hiddens=pos_hiddens, pos_heads=pos_heads,
                                 out_ch=node_embed_dim, heads=heads, activation=F.elu, feat_drop=feat_drop,
                                 attn_drop=attn_drop, negative_slope=negative_slope, residual=res, norm=norm,
                                 p_activation=self.p_act)
        elif mode == "PENL":
            self.gat = GATPSPGNNNL(num_layers=num_gat_layers, in_dim=fv_dim, pos_in_dim=pos_enc_dim,
                                   num_hiddens=num_hiddens, out_ch=node_embed_dim, heads=heads, activation=F.elu,
                                   feat_drop=feat_drop, attn_drop=attn_drop, negative_slope=negative_slope,
                                   residual=res, norm=norm)
        else:
            raise SpgnnError(f"unknown mode {mode!r} (PEL or PENL)")
        self._finish(node_embed_dim, out_ch)

    def forward(self, g, h=None, p=None):
        return self.gat(g, h, p, _head=self.gnn_out)     # (n_out, n_embed, n_p_embed)

    def forward_emb(self, g, h=None, p=None):
        return self.gat(g, h, p)
