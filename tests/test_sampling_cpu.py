"""Neighbour sampling / message-flow blocks (spgnn_b200/sampling.py) — host logic on CPU.

DGL semantics pinned here: per seed min(fanout, in-degree) distinct in-edges; destination nodes are the first source
nodes; blocks chain (layer l's sources are layer l-1's destinations); with fanout = all and every node a seed the
blocks ARE the graph, so the oracle's forward_batch must equal its full-graph forward."""
import numpy as np
import torch

from helpers import FULL_MODELS, golden_graph_inputs


class _HostGraph:
    """in-CSC view of an oracle graph (what sampling.py reads of spgnn_b200.graph.Graph), on CPU."""

    def __init__(self, og):
        self.num_nodes = og.num_nodes
        order = torch.sort(og.dst, stable=True)[1]                     # by destination, ascending edge id inside
        self.in_src = og.src[order].to(torch.int32)
        self.in_ptr = torch.zeros(og.num_nodes + 1, dtype=torch.int32)
        self.in_ptr[1:] = torch.cumsum(torch.bincount(og.dst, minlength=og.num_nodes), 0).to(torch.int32)
        self.ndata = og.ndata


def _graphs():
    from oracle import dgl_ops
    rec, scans = golden_graph_inputs()
    gs = []
    for s in scans:
        g = dgl_ops.graph_from_adj(s["adj"])
        g.ndata["fvs"] = torch.from_numpy(s["fvs"])
        g.ndata["y"] = torch.from_numpy(s["labels"].astype(np.int64))
        gs.append(g)
    og = dgl_ops.batch(gs)
    return og, _HostGraph(og)


def test_sample_neighbors_and_blocks():
    from spgnn_b200 import sampling
    og, hg = _graphs()
    gen = torch.Generator().manual_seed(5)
    edge_set = set(zip(og.src.tolist(), og.dst.tolist()))
    indeg = torch.bincount(og.dst, minlength=og.num_nodes)
    seeds = torch.randperm(og.num_nodes, generator=gen)[:64]
    for fanout in (1, 2, 3, -1):
        src, di = sampling.sample_neighbors(hg, seeds, fanout, gen)
        assert (di[1:] >= di[:-1]).all()                                # grouped by destination
        cnt = torch.bincount(di, minlength=64)
        want = indeg[seeds] if fanout < 0 else indeg[seeds].clamp(max=fanout)
        assert torch.equal(cnt, want)
        pairs = list(zip(src.tolist(), seeds[di].tolist()))
        assert all(p in edge_set for p in pairs) and len(set(pairs)) == len(pairs)     # real edges, no repeats
    # uniform: over many draws every in-edge of a degree-4 node is picked about fanout/deg of the time
    v = int((indeg == 4).nonzero()[0])
    hits = {}
    for _ in range(600):
        src, _ = sampling.sample_neighbors(hg, torch.tensor([v]), 2, gen)
        for u in src.tolist():
            hits[u] = hits.get(u, 0) + 1
    assert len(hits) == 4 and all(abs(h / 600 - 0.5) < 0.08 for h in hits.values())

    blocks = sampling.MultiLayerNeighborSampler([2, 2, 2, 2]).sample_blocks(hg, seeds, gen)
    assert len(blocks) == 4 and torch.equal(blocks[-1].dst_ids, seeds)
    for b, nxt in zip(blocks[:-1], blocks[1:]):
        assert torch.equal(b.dst_ids, nxt.src_ids)                       # layers chain
    for b in blocks:
        assert torch.equal(b.src_ids[:b.num_dst_nodes], b.dst_ids)       # destination i is source i
        assert len(set(b.src_ids.tolist())) == b.num_src_nodes           # every source once
        s_loc, d_loc = b.edges()
        assert all((int(b.src_ids[s]), int(b.dst_ids[d])) in edge_set for s, d in zip(s_loc.tolist(), d_loc.tolist()))
        # out-CSR is the transpose of the in-CSC, out_slot points back into it
        assert int(b.out_ptr[-1]) == b.num_edges == int(b.in_ptr[-1])
        src_of_out = torch.repeat_interleave(torch.arange(b.num_src_nodes), (b.out_ptr[1:] - b.out_ptr[:-1]).long())
        assert torch.equal(b.in_src.long()[b.out_slot.long()], src_of_out)
        assert torch.equal(d_loc[b.out_slot.long()], b.out_dst.long())

    loader = sampling.NodeDataLoader(hg, list(range(0, og.num_nodes, 3)), sampling.MultiLayerNeighborSampler([2, 2]),
                                     batch_size=50, shuffle=True, generator=gen)
    seen = []
    for input_nodes, sd, blks in loader:
        assert torch.equal(blks[0].srcdata["fvs"], og.ndata["fvs"][input_nodes])
        assert torch.equal(blks[-1].dstdata["y"], og.ndata["y"][sd])
        seen += sd.tolist()
    assert sorted(seen) == list(range(0, og.num_nodes, 3)) and len(loader) == -(-len(seen) // 50)


def test_full_fanout_blocks_reproduce_the_full_graph_forward():
    from oracle import dgl_ops, models as om
    from spgnn_b200 import sampling
    og, hg = _graphs()
    kind, cfg = FULL_MODELS["st_sage_3"]
    torch.manual_seed(0)
    net = om.GNNNet(kind, dict(cfg, num_hiddens=[32, 16, 8], node_embed_dim=24, fv_dim=og.ndata["fvs"].shape[1]))
    net.eval()
    blocks = sampling.MultiLayerNeighborSampler([-1] * 4).sample_blocks(hg, torch.arange(og.num_nodes))
    oblocks = []
    for b in blocks:
        s_loc, d_loc = b.edges()
        assert b.num_src_nodes == b.num_dst_nodes == og.num_nodes and b.num_edges == og.src.numel()
        oblocks.append(dgl_ops.Block(s_loc, d_loc, b.num_src_nodes, b.num_dst_nodes))
    with torch.no_grad():
        full = net(og)
        mb = net.forward_batch(oblocks, og.ndata["fvs"])
    for a, r in zip(mb, full):
        assert torch.allclose(a, r, rtol=1e-5, atol=1e-6)
