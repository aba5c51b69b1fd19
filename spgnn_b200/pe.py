"""Positional-encoding inits on device (replace the host networkx code of /root/reference/job_runner.py).

  anchors           job_runner.py:1727-1757 get_anchors_from_cnn_prediction + :1712-1725 add_distal_leafs
  distance_pos_enc  job_runner.py:1759-1777 generate_distant_pos_enc   (the PE the reference actually uses)
  rw_pos_enc        job_runner.py:1684-1702 generate_rw_pos_enc        (dormant in the reference)

Tie rule for distal leaves: the largest node index among the farthest descendant leaves (the reference's pick
depends on CPython set order and is not reproducible; see oracle/pe.py).
"""
from __future__ import annotations

import torch

from ._lib import SpgnnError, lib, ptr, stream
from .ops import empty_padded


def select_anchors(g, fvs_out=None, pos_enc_dim=39):
    """int32 [B, pos_enc_dim] LOCAL node ids (21 CNN anchors + 18 distal leaves when pos_enc_dim == 39)."""
    fo = (g.ndata["fvs_out"] if fvs_out is None else fvs_out).float().contiguous()
    if pos_enc_dim not in (21, 39):
        raise NotImplementedError(f"pos enc dim : {pos_enc_dim}!")
    anchors = torch.empty(g.batch_size, pos_enc_dim, dtype=torch.int32, device=g.device)
    lib().anchor_select(ptr(fo), fo.stride(0), fo.shape[1], ptr(g.node_off), ptr(g.in_ptr), ptr(g.in_src),
                        g.batch_size, pos_enc_dim, g.max_nodes, ptr(anchors), stream())
    return anchors


def distance_pos_enc(g, anchors=None, pos_enc_dim=39, store=True, check=True):
    """pos_enc[n,k] = hops(n, anchor_k) / diameter, fp32 [N, pos_enc_dim]; also returns the per-graph diameters.
    ``check=False`` defers the connectivity check (one device→host read) to :func:`check_connected`."""
    if anchors is None:
        anchors = select_anchors(g, pos_enc_dim=pos_enc_dim)
    anchors = anchors.to(torch.int32).contiguous()
    pos_enc_dim = anchors.shape[1]
    pe = empty_padded(g.num_nodes, pos_enc_dim, g.device)      # 16-byte rows: the projection reads it with 128-bit loads
    diam = torch.empty(g.batch_size, dtype=torch.int32, device=g.device)
    flags = torch.empty(1, dtype=torch.int32, device=g.device)
    nbytes = int(lib().pe_dist_ws_bytes(g.batch_size, g.max_nodes))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=g.device)
    # nx shortest paths run v -> anchor: propagate along out-edges (same graph when the adjacency is symmetric)
    lib().pe_dist_init(ptr(g.node_off), ptr(g.out_ptr), ptr(g.out_dst), ptr(anchors), g.batch_size, pos_enc_dim,
                       g.max_nodes, ptr(pe), pe.stride(0), ptr(diam), ptr(flags), ptr(ws), stream())
    g._pe_flags = flags
    if check:
        check_connected(g)
    if store:
        g.ndata["pos_enc"] = pe
        g.ndata["p"] = pe
    return pe, diam


def check_connected(g):
    """Raises if the last distance_pos_enc of ``g`` met a disconnected graph (nx.diameter raises in the reference)."""
    flags = getattr(g, "_pe_flags", None)
    if flags is not None and int(flags.item()):
        raise SpgnnError("Found infinite path length because the graph is not connected (nx.diameter in the reference)")


def rw_pos_enc(g, pos_enc_dim=39, store=True):
    """diag((A D^-1)^k), k = 1..pos_enc_dim on the self-loop-free graph, fp32 [N, pos_enc_dim]."""
    pe = empty_padded(g.num_nodes, pos_enc_dim, g.device)
    lib().pe_rw_init(ptr(g.node_off), ptr(g.in_ptr), ptr(g.in_src), g.batch_size, pos_enc_dim, g.max_nodes, ptr(pe),
                     pe.stride(0), stream())
    if store:
        g.ndata["rw_enc"] = pe
    return pe
