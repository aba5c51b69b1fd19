"""Neighbour-sampled mini-batches for the SAGE training path of the reference.

Replaces what ``GCNTrainSAGE.train`` (/root/reference/job_runner.py:1460-1514) takes from DGL:
``dgl.dataloading.MultiLayerNeighborSampler(node_ks)``, ``dgl.dataloading.NodeDataLoader(g, nids, sampler,
batch_size=NODE_BATCH_SIZE, shuffle=True)`` and the message-flow-graph ``block`` objects that
``SAGE.forward_batch`` (models.py:685-689) feeds to ``SAGEConv``.

DGL semantics kept (0.7.x ``sample_neighbors(replace=False)`` + ``to_block``):
  * per layer, every seed keeps ``min(fanout, in_degree)`` of its in-edges, drawn uniformly without replacement;
  * a block's destination nodes are the seeds in order, its source nodes are the seeds FIRST and then every other
    endpoint once (so ``h_dst = h_src[:num_dst]``); blocks are built output layer first, each layer's source
    nodes being the previous layer's seeds; ``blocks[0].srcdata`` / ``blocks[-1].dstdata`` carry the node data.
The random stream is torch's, not DGL's: parity with DGL is distributional (which edges are drawn), exact for
everything computed from a given set of blocks.

All index work is torch tensor arithmetic on the graph's device (mini-batches are ~64 seeds x fanout 2 x 4 layers:
latency-bound, no custom kernel); the layer arithmetic on the blocks is this repo's CUDA kernels.
"""
from __future__ import annotations

import torch

from ._lib import SpgnnError


class Block:
    """Bipartite message-flow graph: edges go from ``num_src_nodes`` source rows to ``num_dst_nodes`` destination
    rows; destination node i is source node i.  In-CSC by destination and out-CSR by source (int32), the layout
    the aggregation kernels read (``out_slot`` = position of the edge in the in-CSC)."""

    is_block = True

    def __init__(self, src_local, dst_local, src_ids, num_dst):
        dev = src_local.device
        self.num_src_nodes, self.num_dst_nodes = int(src_ids.numel()), int(num_dst)
        self.src_ids = src_ids                                   # global node ids; [:num_dst] are the seeds
        self.dst_ids = src_ids[:self.num_dst_nodes]
        E = int(src_local.numel())
        # edges arrive grouped by destination in ascending order (the sampler emits them that way)
        if E > 1 and bool((dst_local[1:] < dst_local[:-1]).any()):
            order = torch.sort(dst_local, stable=True)[1]
            src_local, dst_local = src_local[order], dst_local[order]
        self.in_ptr = torch.zeros(self.num_dst_nodes + 1, dtype=torch.int32, device=dev)
        self.in_ptr[1:] = torch.cumsum(torch.bincount(dst_local, minlength=self.num_dst_nodes), 0).to(torch.int32)
        self.in_src = src_local.to(torch.int32)
        order = torch.sort(src_local, stable=True)[1]
        self.out_ptr = torch.zeros(self.num_src_nodes + 1, dtype=torch.int32, device=dev)
        self.out_ptr[1:] = torch.cumsum(torch.bincount(src_local, minlength=self.num_src_nodes), 0).to(torch.int32)
        self.out_dst = dst_local[order].to(torch.int32)
        self.out_slot = order.to(torch.int32)
        self.srcdata, self.dstdata = {}, {}

    # ---- the surface of a DGL block the reference touches (job_runner.py:1499-1501)
    def number_of_src_nodes(self):
        return self.num_src_nodes

    def number_of_dst_nodes(self):
        return self.num_dst_nodes

    def number_of_edges(self):
        return int(self.in_src.numel())

    @property
    def num_edges(self):
        return int(self.in_src.numel())

    @property
    def device(self):
        return self.in_ptr.device

    def int(self):
        return self                                              # indices are int32 already

    def to(self, device=None, **_):
        if device is None or torch.device(device) == self.device:
            return self
        b = object.__new__(Block)
        b.__dict__.update(self.__dict__)
        for k in ("src_ids", "dst_ids", "in_ptr", "in_src", "out_ptr", "out_dst", "out_slot"):
            setattr(b, k, getattr(self, k).to(device))
        b.srcdata = {k: v.to(device) for k, v in self.srcdata.items()}
        b.dstdata = {k: v.to(device) for k, v in self.dstdata.items()}
        return b

    def check_no_zero_in_degree(self):
        if bool((self.in_ptr[1:] == self.in_ptr[:-1]).any()):
            raise SpgnnError("There are 0-in-degree nodes in the block")

    def edges(self):
        """(src_local, dst_local) int64 in in-CSC order."""
        deg = (self.in_ptr[1:] - self.in_ptr[:-1]).long()
        dst = torch.repeat_interleave(torch.arange(self.num_dst_nodes, device=self.device), deg)
        return self.in_src.long(), dst


def sample_neighbors(g, seeds, fanout, generator=None):
    """For every seed keep min(fanout, in_degree) in-edges, uniformly without replacement (fanout < 0: all).
    Returns (src global ids, dst index into ``seeds``), grouped by destination in ascending order."""
    dev = g.in_ptr.device
    seeds = seeds.to(dev).long()
    S = int(seeds.numel())
    beg = g.in_ptr[seeds].long()
    deg = g.in_ptr[seeds + 1].long() - beg
    total = int(deg.sum())
    seg = torch.repeat_interleave(torch.arange(S, device=dev), deg)
    start = torch.cumsum(deg, 0) - deg                           # first edge of each seed in the enumeration
    off = torch.arange(total, device=dev) - start[seg]
    slot = beg[seg] + off
    if fanout >= 0 and total > 0:
        key = torch.rand(total, generator=generator, device=dev, dtype=torch.float64)
        order = torch.sort(seg.to(torch.float64) + key)[1]       # random order inside every seed's segment
        rank = torch.arange(total, device=dev) - start[seg[order]]
        keep = order[rank < fanout]
        keep = torch.sort(keep)[0]                               # back to ascending (destination, edge id) order
        seg, slot = seg[keep], slot[keep]
    return g.in_src[slot].long(), seg


def to_block(g, src, dst_index, seeds):
    """DGL ``to_block``: destination nodes = seeds (in order), source nodes = seeds first, then the other endpoints."""
    dev = src.device
    seeds = seeds.to(dev).long()
    S = int(seeds.numel())
    pos = torch.full((g.num_nodes,), -1, dtype=torch.int64, device=dev)
    pos[seeds] = torch.arange(S, device=dev)
    extra = torch.unique(src[pos[src] < 0])
    pos[extra] = S + torch.arange(int(extra.numel()), device=dev)
    return Block(pos[src], dst_index, torch.cat([seeds, extra]), S)


class MultiLayerNeighborSampler:
    """``dgl.dataloading.MultiLayerNeighborSampler(fanouts)``: one block per layer, ``fanouts[i]`` for layer i."""

    def __init__(self, fanouts, replace=False, return_eids=False):
        if replace:
            raise SpgnnError("MultiLayerNeighborSampler: only replace=False (what the reference uses)")
        self.fanouts = [int(f) if f is not None else -1 for f in fanouts]

    def sample_blocks(self, g, seeds, generator=None):
        blocks = []
        seeds = torch.as_tensor(seeds, device=g.in_ptr.device).long()
        for fanout in reversed(self.fanouts):
            src, dst_index = sample_neighbors(g, seeds, fanout, generator)
            block = to_block(g, src, dst_index, seeds)
            blocks.insert(0, block)
            seeds = block.src_ids
        return blocks


class NodeDataLoader:
    """``dgl.dataloading.NodeDataLoader(g, nids, sampler, batch_size=, shuffle=, drop_last=)``: iterates
    ``(input_nodes, seeds, blocks)`` with ``blocks[0].srcdata`` / ``blocks[-1].dstdata`` filled from ``g.ndata``."""

    def __init__(self, g, nids, block_sampler, device=None, batch_size=1, shuffle=False, drop_last=False,
                 num_workers=0, generator=None, **_):
        self.g, self.sampler = g, block_sampler
        self.nids = torch.as_tensor(nids, device=g.in_ptr.device).long()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator

    def __len__(self):
        n = int(self.nids.numel())
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        nids = self.nids
        if self.shuffle:
            perm = torch.randperm(int(nids.numel()), generator=self.generator, device=nids.device)
            nids = nids[perm]
        for i in range(len(self)):
            seeds = nids[i * self.batch_size:(i + 1) * self.batch_size]
            blocks = self.sampler.sample_blocks(self.g, seeds, self.generator)
            input_nodes = blocks[0].src_ids
            for k, v in self.g.ndata.items():
                blocks[0].srcdata[k] = v[input_nodes]
                blocks[-1].dstdata[k] = v[seeds]
            yield input_nodes, seeds, blocks
