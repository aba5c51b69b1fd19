"""Shared helpers for the test-suite (settings dicts, golden loaders, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tests/golden/make_golden.py:TINY — widths of the committed wiring fixtures
TINY_COMMON = dict(out_ch=22, fv_dim=32, num_hiddens=[16, 8, 8], node_embed_dim=24)
TINY_MODELS = {
    "gat3": ("gat", dict(TINY_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1, attn_drop=0.1,
                         negative_slope=0.2)),
    "gat3_nr": ("gat", dict(TINY_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                            attn_drop=0.1, negative_slope=0.2, res=False)),
    "gat6": ("gat", dict(TINY_COMMON, num_gat_layers=6, num_heads=2, num_out_heads=2, feat_drop=0.1, attn_drop=0.1,
                         negative_slope=0.2, num_hiddens=[16, 8, 8, 8, 8, 8])),
    "gcn3": ("gcn", dict(TINY_COMMON, num_gcn_layers=3)),
    "gin3": ("gin", dict(TINY_COMMON, num_gin_layers=3)),
    "sage3": ("sage", dict(TINY_COMMON, num_layers=3, feat_drop=0.1, node_ks=[2, 2, 2, 2], node_sample_rate=0.3)),
    "spgnn3": ("spgnn", dict(TINY_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                             attn_drop=0.1, negative_slope=0.2, pos_hiddens=[16, 8, 8], num_pos_heads=1,
                             pos_enc_dim=39)),
    "spgnnnl3": ("spgnn", dict(TINY_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                               attn_drop=0.1, negative_slope=0.2, pos_hiddens=[16, 8, 8], num_pos_heads=1,
                               pos_enc_dim=39, mode="PENL")),
}

# exp_settings/*.py MODEL dicts, GNN keys only
FULL_COMMON = dict(out_ch=22, fv_dim=1024, num_hiddens=[256, 128, 64], node_embed_dim=1024)
FULL_MODELS = {
    "st_gat_3": ("gat", dict(FULL_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                             attn_drop=0.1, negative_slope=0.2)),
    "st_gat_6": ("gat", dict(FULL_COMMON, num_gat_layers=6, num_heads=2, num_out_heads=2, feat_drop=0.1,
                             attn_drop=0.1, negative_slope=0.2, num_hiddens=[256, 128, 64, 64, 64, 64])),
    "st_gat_6_nr": ("gat", dict(FULL_COMMON, num_gat_layers=6, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                attn_drop=0.1, negative_slope=0.2, num_hiddens=[256, 128, 64, 64, 64, 64],
                                res=False)),
    "st_gcn_3": ("gcn", dict(FULL_COMMON, num_gcn_layers=3)),
    "st_gin_3": ("gin", dict(FULL_COMMON, num_gin_layers=3)),
    "st_sage_3": ("sage", dict(FULL_COMMON, num_layers=3, feat_drop=0.1, node_ks=[2, 2, 2, 2],
                               node_sample_rate=0.3)),
    "st_pgat_spgnn_3": ("spgnn", dict(FULL_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                      attn_drop=0.1, negative_slope=0.2, pos_hiddens=[256, 128, 64],
                                      num_pos_heads=1, pos_enc_dim=39)),
    "st_pgat_spgnnnl_3": ("spgnn", dict(FULL_COMMON, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                        attn_drop=0.1, negative_slope=0.2, pos_hiddens=[256, 128, 64],
                                        num_pos_heads=1, pos_enc_dim=39, mode="PENL")),
}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def golden_state_dict(rec):
    return {k[4:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("sd::")}


def golden_graph_inputs():
    """The 4-tree batch of tests/golden/graph_pe.npz as python lists."""
    rec = load_golden("graph_pe.npz")
    n = int(rec["n_graphs"])
    scans = [dict(adj=rec[f"adj{i}"], fvs=rec[f"fvs{i}"], fvs_out=rec[f"fvs_out{i}"], labels=rec[f"labels{i}"])
             for i in range(n)]
    return rec, scans


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
