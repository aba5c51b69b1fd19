"""Synthetic airway-tree batches generated ON DEVICE (bench input).

Device twin of ``spgnn_b200.synth``: the integer part (tree shapes, labels, DGL edge order) is bit-identical to
the host generator for the same (seed, tree index); normals agree to fp32 rounding.  Batches are a pure
function of (seed, first_tree, count), so shards are reproducible regardless of the GPU count.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ._lib import lib, ptr, stream
from .graph import Graph, device_scan


@dataclass
class DeviceBatch:
    graph: Graph
    parent: torch.Tensor      # int64 [N] local parent ids (-1 for roots)


def make_batch(first_tree, count, seed=1234, ragged=False, k_fixed=150, fv_dim=1024, n_class=22, features=True,
               device=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = lib()
    n_nodes = torch.empty(count, dtype=torch.int64, device=dev)
    n_edges = torch.empty(count, dtype=torch.int64, device=dev)
    L.synth_sizes(first_tree, count, seed, int(ragged), k_fixed, ptr(n_nodes), ptr(n_edges), stream())
    node_off, edge_off = device_scan(n_nodes), device_scan(n_edges)
    N, E = int(node_off[-1].item()), int(edge_off[-1].item())
    parent = torch.empty(N, dtype=torch.int64, device=dev)
    labels = torch.empty(N, dtype=torch.int64, device=dev)
    L.synth_trees(first_tree, count, seed, int(ragged), k_fixed, ptr(node_off), ptr(parent), ptr(labels), stream())
    sl = torch.empty(E, dtype=torch.int64, device=dev)
    dl = torch.empty(E, dtype=torch.int64, device=dev)
    L.synth_edges(ptr(node_off), ptr(edge_off), ptr(parent), count, ptr(sl), ptr(dl), stream())
    max_nodes = 2 * (180 if ragged else k_fixed) + 1
    g = Graph.from_edge_lists(n_nodes, n_edges, sl, dl, max_nodes=max_nodes, check=False)
    g.ndata["y"] = labels
    if features:
        fvs = torch.empty(N, fv_dim, dtype=torch.float32, device=dev)
        fo = torch.empty(N, n_class, dtype=torch.float32, device=dev)
        L.synth_features(first_tree, count, seed, ptr(node_off), N, ptr(fvs), fvs.stride(0), fv_dim, ptr(fo),
                         fo.stride(0), n_class, stream())
        g.ndata["fvs"], g.ndata["fvs_out"] = fvs, fo
    return DeviceBatch(g, parent)
