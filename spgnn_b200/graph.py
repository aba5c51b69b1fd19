"""Device-resident batched graph: the DGLGraph / dgl.batch subset the reference uses, built by CUDA kernels.

Replaces: ``DGLGraph(nx_graph)``, ``dgl.remove_self_loop``, ``g.add_edges``, ``g.nodes()``,
``g.ndata``, ``g.number_of_nodes()``, ``dgl.batch``, ``dgl.unbatch``, ``g.batch_size``, ``g.in_degrees()``,
``g.adjacency_matrix()``, ``g.to()`` as called from /root/reference/job_runner.py:822-838, :1319-1344,
:1779-1801, :1390, :1882, :2046.  Edge ids, node offsets and edge order are bit-identical to DGL's; on top of
that the builder emits the in-CSC / out-CSR the aggregation kernels read.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import SpgnnError, lib, ptr, stream


def _dev(device=None):
    if not torch.cuda.is_available():
        raise SpgnnError("spgnn_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def device_scan(counts: torch.Tensor) -> torch.Tensor:
    """Exclusive prefix scan on device: int64 [n] → int64 [n+1] (``out[n]`` = total)."""
    counts = counts.contiguous()
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    ws = torch.empty(max(int(lib().scan_ws_bytes(n)), 8), dtype=torch.uint8, device=counts.device)
    lib().scan_i64(ptr(counts), ptr(out), n, ptr(ws), stream())
    return out


class Graph:
    """A batch of graphs on one GPU.  Build with :func:`from_adj`, :func:`batch`, :func:`batch_from_adjs`
    or :meth:`Graph.from_edge_lists`."""

    def __init__(self):
        self.ndata = {}
        self._norms = None

    # ---------------------------------------------------------------- construction
    @classmethod
    def from_edge_lists(cls, n_nodes, n_edges, src_local, dst_local, max_nodes=None, check=True, num_nodes=None,
                        max_degree=None, zero_in_degree=None):
        """``dgl.batch`` of B graphs given per-graph node/edge counts (int64 [B]) and the concatenated LOCAL
        edge lists in DGL edge order (int64 [E]); all device tensors.  ``num_nodes`` / ``max_degree`` /
        ``zero_in_degree``: values the caller already knows on the host (a loader that packed the batch): with
        them and ``check=False`` the build issues no device→host read, so a pipelined caller never stalls."""
        g = cls()
        dev = n_nodes.device
        g.device = dev
        g._bnn = n_nodes.contiguous()
        g._bne = n_edges.contiguous()
        B = g._bnn.numel()
        g.node_off = device_scan(g._bnn)
        g.edge_off = device_scan(g._bne)
        N = int(g.node_off[-1].item()) if num_nodes is None else int(num_nodes)
        E = int(src_local.shape[0])
        if check and int(g.edge_off[-1].item()) != E:
            raise SpgnnError("edge list length does not match the per-graph edge counts")
        g.num_nodes, g.num_edges, g.batch_size = N, E, B
        g.src_local, g.dst_local = src_local.contiguous(), dst_local.contiguous()
        i64 = dict(dtype=torch.int64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        g.src, g.dst = torch.empty(E, **i64), torch.empty(E, **i64)
        g.node_gid = torch.empty(N, **i32)
        g.in_ptr, g.in_src, g.in_eid = torch.empty(N + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
        g.out_ptr, g.out_dst, g.out_slot = torch.empty(N + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
        flags = torch.empty(4, **i32)
        ws = torch.empty(int(lib().batch_ws_bytes(N, E)), dtype=torch.uint8, device=dev)
        lib().batch_build(ptr(g.node_off), ptr(g.edge_off), B, N, E, ptr(g.src_local), ptr(g.dst_local),
                          ptr(g.src), ptr(g.dst), ptr(g.node_gid), ptr(g.in_ptr), ptr(g.in_src), ptr(g.in_eid),
                          ptr(g.out_ptr), ptr(g.out_dst), ptr(g.out_slot), ptr(flags), ptr(ws), stream())
        g._flags = flags
        g.zero_in_degree, g._max_degree = zero_in_degree, max_degree
        if check:
            f = g._read_flags()
            if f[1]:
                raise SpgnnError(f"{f[1]} edge endpoints are outside their graph's node range")
        g.max_nodes = int(g._bnn.max().item()) if max_nodes is None else int(max_nodes)
        return g

    # ---------------------------------------------------------------- DGLGraph surface
    def number_of_nodes(self):
        return self.num_nodes

    def number_of_edges(self):
        return self.num_edges

    def nodes(self):
        return torch.arange(self.num_nodes, dtype=torch.int64, device=self.device)

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    def in_degrees(self):
        return (self.in_ptr[1:] - self.in_ptr[:-1]).to(torch.int64)

    def out_degrees(self):
        return (self.out_ptr[1:] - self.out_ptr[:-1]).to(torch.int64)

    def to(self, device=None, **_):
        if device is not None and torch.device(device).type != "cuda":
            raise SpgnnError("spgnn_b200 graphs live on CUDA devices only")
        return self

    def cpu(self):
        """``g.cpu()``: a read-only host copy (:class:`HostGraph`) for the host-side helpers the reference runs on it —
        ``dgl.to_networkx(g.cpu())`` (job_runner.py:1763), ``g.cpu().to_networkx()`` (:1652),
        ``adjacency_matrix().to_dense().numpy()`` (:1742).  The kernels never take it: ``.to('cuda')`` gives back the
        device graph it was copied from."""
        return HostGraph(self)

    def local_var(self):
        return self

    def add_edges(self, u, v):
        """``g.add_edges(u, v)`` for an unbatched graph (job_runner.py:1800: self loops appended last)."""
        if self.batch_size != 1:
            raise SpgnnError("add_edges is only supported on an unbatched graph")
        u = torch.as_tensor(u, dtype=torch.int64, device=self.device)
        v = torch.as_tensor(v, dtype=torch.int64, device=self.device)
        new = Graph.from_edge_lists(self._bnn, self._bne + u.numel(), torch.cat([self.src_local, u]),
                                    torch.cat([self.dst_local, v]))
        nd = self.ndata
        self.__dict__.update(new.__dict__)
        self.ndata = nd

    def adjacency_matrix(self, transpose=False, scipy_fmt=None):
        """DGL 0.7 ``g.adjacency_matrix()`` (dense device tensor, A[src, dst]); with ``scipy_fmt`` ("csr" / "coo") a
        host scipy matrix, as job_runner.py:1814 asks for."""
        if scipy_fmt is not None:
            return self.adjacency_matrix_scipy(transpose=transpose, fmt=scipy_fmt)
        a = torch.zeros(self.num_nodes, self.num_nodes, device=self.device)
        a[self.src, self.dst] = 1.0
        return a.t() if transpose else a

    def adjacency_matrix_scipy(self, transpose=False, fmt="csr", return_edge_ids=False):
        """``g.adjacency_matrix_scipy(return_edge_ids=False)`` (job_runner.py:1632): host scipy sparse matrix with
        ones (or the edge ids) at [src, dst]."""
        import scipy.sparse as sp
        s, d = self.src.cpu().numpy(), self.dst.cpu().numpy()
        if transpose:
            s, d = d, s
        vals = np.arange(self.num_edges) if return_edge_ids else np.ones(self.num_edges)
        m = sp.coo_matrix((vals, (s, d)), shape=(self.num_nodes, self.num_nodes))
        return m.asformat(fmt)

    def to_networkx(self):
        """``dgl.to_networkx(g)`` / ``g.to_networkx()`` (job_runner.py:1763): a host nx.MultiDiGraph with the edges in
        id order and an ``id`` attribute per edge."""
        import networkx as nx
        G = nx.MultiDiGraph()
        G.add_nodes_from(range(self.num_nodes))
        for i, (u, v) in enumerate(zip(self.src.cpu().tolist(), self.dst.cpu().tolist())):
            G.add_edge(u, v, id=i)
        return G

    def norms(self):
        """(outdeg^-1/2, indeg^-1/2, 1/indeg, 1/outdeg), degrees clamped at 1 (GraphConv 'both', GIN 'mean')."""
        if self._norms is None:
            N = self.num_nodes
            f = dict(dtype=torch.float32, device=self.device)
            o_s, i_s, i_i, o_i = (torch.empty(N, **f) for _ in range(4))
            lib().degree_norms(ptr(self.out_ptr), N, ptr(o_s), ptr(o_i), stream())
            lib().degree_norms(ptr(self.in_ptr), N, ptr(i_s), ptr(i_i), stream())
            self._norms = (o_s, i_s, i_i, o_i)
        return self._norms

    def _read_flags(self):
        """One device→host read of the batch builder's flags (zero in-degree count, bad endpoints, max degrees)."""
        f = self._flags.tolist()
        self.zero_in_degree, self._max_degree = f[0], max(f[2], f[3])
        return f

    def max_degree(self):
        """Largest in- or out-degree of the batch (airway trees with self loops: 4)."""
        if self._max_degree is None:
            self._read_flags()
        return self._max_degree

    def check_no_zero_in_degree(self):
        if self.zero_in_degree is None:
            self._read_flags()
        if self.zero_in_degree:
            raise SpgnnError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid "
                             "(DGLError in the reference stack); add self loops or pass allow_zero_in_degree=True")


class HostGraph:
    """Host copy of a device graph's structure (edge lists in edge-id order) and node data: what ``Graph.cpu()`` returns.
    Read-only DGLGraph subset — sizes, ``nodes`` / ``edges``, degrees, adjacency (dense tensor, scipy), networkx."""

    def __init__(self, g):
        self._device_graph = g
        self.src, self.dst = g.src.cpu(), g.dst.cpu()
        self.num_nodes, self.num_edges = int(g.num_nodes), int(self.src.numel())
        self.batch_size = g.batch_size
        self._bnn, self._bne = g._bnn.cpu(), g._bne.cpu()
        self.ndata = {k: v.cpu() for k, v in g.ndata.items()}
        self.device = torch.device("cpu")

    def number_of_nodes(self):
        return self.num_nodes

    def number_of_edges(self):
        return self.num_edges

    def nodes(self):
        return torch.arange(self.num_nodes, dtype=torch.int64)

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.num_nodes)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.num_nodes)

    def to(self, device=None, **_):
        if device is None or torch.device(device).type == "cpu":
            return self
        return self._device_graph.to(device)

    def cpu(self):
        return self

    def adjacency_matrix(self, transpose=False, scipy_fmt=None):
        if scipy_fmt is not None:
            return self.adjacency_matrix_scipy(transpose=transpose, fmt=scipy_fmt)
        a = torch.zeros(self.num_nodes, self.num_nodes)
        a[self.src, self.dst] = 1.0
        return a.t() if transpose else a

    def adjacency_matrix_scipy(self, transpose=False, fmt="csr", return_edge_ids=False):
        import scipy.sparse as sp
        s, d = self.src.numpy(), self.dst.numpy()
        if transpose:
            s, d = d, s
        vals = np.arange(self.num_edges) if return_edge_ids else np.ones(self.num_edges)
        return sp.coo_matrix((vals, (s, d)), shape=(self.num_nodes, self.num_nodes)).asformat(fmt)

    def to_networkx(self):
        import networkx as nx
        G = nx.MultiDiGraph()
        G.add_nodes_from(range(self.num_nodes))
        for i, (u, v) in enumerate(zip(self.src.tolist(), self.dst.tolist())):
            G.add_edge(u, v, id=i)
        return G


# -------------------------------------------------------------------- builders
def _edges_from_dense(adj_cat, n_nodes, dev):
    """Device dense adj → DGL edge list per graph (off-diagonal non-zeros row-major, then self loops)."""
    B = n_nodes.numel()
    node_off = device_scan(n_nodes)
    adj_off = device_scan(n_nodes * n_nodes)
    N = int(node_off[-1].item())
    row_cnt = torch.empty(N, dtype=torch.int64, device=dev)
    n_edges = torch.empty(B, dtype=torch.int64, device=dev)
    L = lib()
    L.adj_count(ptr(adj_cat), ptr(adj_off), ptr(n_nodes), ptr(node_off), B, N, ptr(row_cnt), ptr(n_edges), stream())
    edge_off = device_scan(n_edges)
    row_off = device_scan(row_cnt)
    E = int(edge_off[-1].item())
    sl = torch.empty(E, dtype=torch.int64, device=dev)
    dl = torch.empty(E, dtype=torch.int64, device=dev)
    L.adj_fill(ptr(adj_cat), ptr(adj_off), ptr(n_nodes), ptr(node_off), ptr(edge_off), ptr(row_off), B, N,
               ptr(sl), ptr(dl), stream())
    return n_edges, sl, dl


def batch_from_adjs(adjs, device=None):
    """``dgl.batch([from_adj_to_graph(a) for a in adjs])`` in one pass: dense uint8 adjacencies (tree ∪ I,
    numpy or tensors) → one H2D copy → device COO + CSC build (job_runner.py:1779-1801 + :1882)."""
    dev = _dev(device)
    if isinstance(adjs[0], torch.Tensor) and adjs[0].is_cuda:
        n_list = [int(a.shape[0]) for a in adjs]
        adj_cat = torch.cat([(a != 0).to(torch.uint8).reshape(-1) for a in adjs])
    else:
        arrs = [np.ascontiguousarray((np.asarray(a) != 0).astype(np.uint8)) for a in adjs]
        n_list = [a.shape[0] for a in arrs]
        host = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs]))
        adj_cat = host.pin_memory().to(dev, non_blocking=True)
    for a, n in zip(adjs, n_list):
        if a.shape[0] != a.shape[1]:
            raise SpgnnError("adjacency must be square")
    n_nodes = torch.tensor(n_list, dtype=torch.int64).to(dev)
    n_edges, sl, dl = _edges_from_dense(adj_cat, n_nodes, dev)
    return Graph.from_edge_lists(n_nodes, n_edges, sl, dl, max_nodes=max(n_list))


def from_adj(adj, device=None):
    """One graph from a dense adjacency — ``from_adj_to_graph`` of the reference (self loops appended last)."""
    return batch_from_adjs([adj], device)


def from_edges(src, dst, num_nodes, device=None):
    """One graph from explicit edge lists (``DGLGraph`` + ``add_edges``)."""
    dev = _dev(device)
    s = torch.as_tensor(src, dtype=torch.int64).to(dev)
    d = torch.as_tensor(dst, dtype=torch.int64).to(dev)
    return Graph.from_edge_lists(torch.tensor([num_nodes], dtype=torch.int64, device=dev),
                                 torch.tensor([s.numel()], dtype=torch.int64, device=dev), s, d, max_nodes=num_nodes)


def DGLGraph(data=None, device=None):
    """``dgl.DGLGraph(nx_graph)`` as the reference calls it (job_runner.py:1781-1783: ``DGLGraph(nx.DiGraph(adj))``):
    nodes 0..n-1 in sorted order, edges in ``nx_graph.edges()`` order (DGL's from_networkx for a graph without edge
    ids); an undirected graph contributes both directions.  A dense adjacency (numpy / tensor) is accepted as well and
    means ``DGLGraph(nx.DiGraph(adj))`` — non-zero entries row-major, self loops where the diagonal has them."""
    if data is None:
        raise SpgnnError("DGLGraph(): an empty graph cannot be grown node by node here; pass a networkx graph or an adjacency")
    if hasattr(data, "edges") and hasattr(data, "number_of_nodes"):          # networkx Graph / DiGraph
        if not data.is_directed():
            data = data.to_directed()
        nodes = sorted(data.nodes())
        index = {v: i for i, v in enumerate(nodes)}
        e = np.asarray([(index[u], index[v]) for u, v in data.edges()], dtype=np.int64).reshape(-1, 2)
        return from_edges(e[:, 0], e[:, 1], len(nodes), device)
    adj = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    if adj.ndim != 2 or adj.shape[0] != adj.shape[1]:
        raise SpgnnError("adjacency must be square")
    src, dst = np.nonzero(adj)                                                # row-major: nx.DiGraph(adj).edges() order
    return from_edges(src.astype(np.int64), dst.astype(np.int64), adj.shape[0], device)


def to_networkx(g):
    return g.to_networkx()


def remove_self_loop(g):
    if g.batch_size != 1:
        raise SpgnnError("remove_self_loop is only supported on an unbatched graph")
    keep = g.src_local != g.dst_local
    out = from_edges(g.src_local[keep], g.dst_local[keep], g.num_nodes, g.device)
    out.ndata = dict(g.ndata)
    return out


def batch(graphs):
    """``dgl.batch``: disjoint union, node ids shifted by Σ n_j, per-graph edge order kept, ndata concatenated."""
    if len(graphs) == 0:
        raise SpgnnError("batch() of an empty list")
    n_nodes = torch.cat([g._bnn for g in graphs])
    n_edges = torch.cat([g._bne for g in graphs])
    # a member that is itself a batch contributes ids relative to its own graphs: src_local/dst_local already are
    sl = torch.cat([g.src_local for g in graphs])
    dl = torch.cat([g.dst_local for g in graphs])
    out = Graph.from_edge_lists(n_nodes, n_edges, sl, dl, max_nodes=max(g.max_nodes for g in graphs))
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    return out


def unbatch(g):
    outs = []
    no, eo = g.node_off.tolist(), g.edge_off.tolist()
    for i in range(g.batch_size):
        h = from_edges(g.src_local[eo[i]:eo[i + 1]], g.dst_local[eo[i]:eo[i + 1]], no[i + 1] - no[i], g.device)
        for k, v in g.ndata.items():
            h.ndata[k] = v[no[i]:no[i + 1]]
        outs.append(h)
    return outs
