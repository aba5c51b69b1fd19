"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN CODE (authoring container only).

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.npz

What runs from /root/reference, unmodified:
  * ``models.py`` — imported as a module (classes GAT, GCN, GIN, SAGE, GATPSPGNN,
    GATPSPGNNNL and the ``*Net`` wrappers).  Its two missing third-party imports
    are satisfied by shims: ``dgl`` / ``dgl.nn.pytorch`` → ``oracle.dgl_ops`` (the
    DGL-0.7.x restatement; DGL itself cannot be installed here) and ``utils`` →
    two unused names.  This pins the WIRING (layer widths, drop placement, concat
    order, head flatten/mean, return tuples), not DGL's arithmetic.
  * ``job_runner.py`` — the bodies of ``GCNTrainSPGNN.from_adj_to_graph``,
    ``generate_distant_pos_enc``, ``get_anchors_from_cnn_prediction``,
    ``add_distal_leafs``, ``generate_rw_pos_enc`` (:1684-1801) and
    ``GCNTest.from_adj_to_graph`` (:822-838), ``JobRunner._prediction_by_branch_probs``
    (:158-165 arg-max line) are extracted with ``ast`` and executed against a graph
    shim (job_runner.py itself cannot be imported: SimpleITK, skimage, tensorboardX,
    seaborn, matplotlib, dgl are absent).  networkx is real.  ``np.float`` (removed
    in numpy 2) and ``Tensor.cuda()`` (no GPU here) are aliased.

The fixtures are small (tiny widths, trees of 23-41 nodes) and are committed;
tests/test_oracle_golden.py checks the oracle against them, tests/test_gpu_*.py
check the CUDA path against them.
"""
from __future__ import annotations

import ast
import os
import sys
import types
import warnings

import networkx as nx
import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import dgl_ops  # noqa: E402
from spgnn_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")


# ----------------------------------------------------------------------------
# shims
# ----------------------------------------------------------------------------
class _Adj:
    def __init__(self, dense):
        self._d = dense

    def to_dense(self):
        return self._d


class ShimGraph(dgl_ops.Graph):
    """oracle Graph + the handful of DGLGraph methods job_runner.py:1684-1801 calls."""

    def adjacency_matrix(self, *a, **k):
        return _Adj(self.adjacency_dense())


def DGLGraph(nx_graph):
    """``DGLGraph(nx_graph)`` [DGL-upstream]: nodes 0..n-1, edges in ``nx_graph.edges`` iteration order
    (an undirected nx.Graph is first converted with ``to_directed()``)."""
    if not nx_graph.is_directed():
        nx_graph = nx_graph.to_directed()
    e = list(nx_graph.edges())
    src = np.array([u for u, _ in e], dtype=np.int64)
    dst = np.array([v for _, v in e], dtype=np.int64)
    return ShimGraph(src, dst, nx_graph.number_of_nodes())


def _remove_self_loop(g):
    keep = g.src != g.dst
    out = ShimGraph(g.src[keep], g.dst[keep], g.num_nodes)
    out.ndata = dict(g.ndata)
    return out


def _to_networkx(g):
    G = nx.MultiDiGraph()
    G.add_nodes_from(range(g.num_nodes))
    G.add_edges_from(zip(g.src.tolist(), g.dst.tolist()))
    return G


def install_shims():
    dgl = types.ModuleType("dgl")
    dgl.DGLGraph = DGLGraph
    dgl.batch = dgl_ops.batch
    dgl.unbatch = dgl_ops.unbatch
    dgl.remove_self_loop = _remove_self_loop
    dgl.to_networkx = _to_networkx
    dgl.backend = types.SimpleNamespace(asnumpy=lambda t: t.numpy())
    dgl_nn = types.ModuleType("dgl.nn")
    dgl_nn_pt = types.ModuleType("dgl.nn.pytorch")
    for name in ("GATConv", "GraphConv", "SAGEConv", "GINConv"):
        setattr(dgl_nn_pt, name, getattr(dgl_ops, name))
    dgl_nn_pt.AvgPooling = dgl_nn_pt.MaxPooling = object        # imported at models.py:8, never used
    dgl.nn = dgl_nn
    dgl_nn.pytorch = dgl_nn_pt
    utils = types.ModuleType("utils")
    utils.topk = utils.get_batch_id = None                        # models.py:7; only used by dead code
    sys.modules.update({"dgl": dgl, "dgl.nn": dgl_nn, "dgl.nn.pytorch": dgl_nn_pt, "utils": utils})
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not hasattr(np, "float"):
        np.float = float
    return dgl


def load_runner_methods(dgl):
    """Extract method bodies from job_runner.py by name and bind them to a bare class."""
    tree = ast.parse(open(os.path.join(REF, "job_runner.py")).read())
    want = {
        "GCNTrainSPGNN": ["from_adj_to_graph", "generate_distant_pos_enc", "get_anchors_from_cnn_prediction",
                          "add_distal_leafs", "generate_rw_pos_enc"],
        "GCNTest": ["from_adj_to_graph"],
        "JobRunner": ["_prediction_by_branch_probs"],
    }
    ns = {"np": np, "nx": nx, "torch": torch, "F": F, "dgl": dgl, "DGLGraph": DGLGraph,
          "visualize_airway_graph": lambda *a, **k: None}
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in want:
            fns = [f for f in node.body if isinstance(f, ast.FunctionDef) and f.name in want[node.name]]
            mod = ast.Module(body=[ast.ClassDef(name=node.name, bases=[], keywords=[], body=fns, decorator_list=[])],
                             type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, f"{REF}/job_runner.py", "exec"), ns)
            out[node.name] = ns[node.name]
    return out


class _Log:
    def debug(self, *a, **k):
        pass

    info = debug


def make_runner(cls, pos_enc_dim=39):
    r = cls.__new__(cls)
    r.settings = types.SimpleNamespace(POS_ENC_DIM=pos_enc_dim, GRAPH_MODE="all_connected")
    r.logger = _Log()
    r.trace = False
    return r


# ----------------------------------------------------------------------------
# fixtures
# ----------------------------------------------------------------------------
TINY = dict(fv_dim=32, num_hiddens=[16, 8, 8], pos_hiddens=[16, 8, 8], node_embed_dim=24, pos_enc_dim=39, out_ch=22)
CNN_STUB = dict(n_layers=1, in_ch_list=[1, 2], base_ch_list=[2, 2], end_ch_list=[2, 2], checkpoint_layers=[0, 0],
                kernel_sizes=[1, 1], padding_list=[(0, 0, 0), (0, 0, 0)], conv_strides=[[1, 1]], dropout=0.0,
                spatial_size=1)
TREE_KS = [11, 14, 20, 12]     # n = 23, 29, 41, 25


def tiny_scans(seed=7):
    return [synth.make_scan(100 + i, seed=seed, fv_dim=TINY["fv_dim"], k=k) for i, k in enumerate(TREE_KS)]


def gen_graph_and_pe(R, out):
    scans = tiny_scans()
    spg = make_runner(R["GCNTrainSPGNN"])
    tst = make_runner(R["GCNTest"])
    g_list, rec = [], {}
    for i, sc in enumerate(scans):
        adj = torch.from_numpy(sc.adj).float()
        # --- reference: SPGNN construction + live distance PE (job_runner.py:1779-1801) ---
        g = spg.from_adj_to_graph(adj, torch.from_numpy(sc.fvs), torch.from_numpy(sc.fvs_out),
                                  torch.from_numpy(sc.labels), None, i, f"u{i}")
        rec[f"src{i}"], rec[f"dst{i}"] = g.src.numpy(), g.dst.numpy()
        rec[f"pos_enc{i}"] = g.ndata["pos_enc"].numpy()
        # anchors again (deterministic given fvs_out) to store them
        g0 = _remove_self_loop(g)
        g0.ndata = dict(g.ndata)
        rec[f"anchors{i}"] = np.asarray(spg.get_anchors_from_cnn_prediction(g0, i, f"u{i}"), dtype=np.int64)
        rec[f"all_pos{i}"] = spg.generate_distant_pos_enc(g0, i, f"u{i}").numpy()
        # --- reference: dormant RW PE on the self-loop-free graph (call site commented out at :1792) ---
        spg.generate_rw_pos_enc(g0)
        rec[f"rw_enc{i}"] = g0.ndata["rw_enc"].numpy()
        # --- reference: GCNTest construction path (:822-838), symmetric adj → nx.Graph ---
        g2 = tst.from_adj_to_graph(adj)
        rec[f"src_sym{i}"], rec[f"dst_sym{i}"] = g2.src.numpy(), g2.dst.numpy()
        rec[f"adj{i}"], rec[f"fvs{i}"], rec[f"fvs_out{i}"], rec[f"labels{i}"] = sc.adj, sc.fvs, sc.fvs_out, sc.labels
        g_list.append(g)
    bg = sys.modules["dgl"].batch(g_list)
    rec["b_src"], rec["b_dst"] = bg.src.numpy(), bg.dst.numpy()
    rec["b_num_nodes"], rec["b_num_edges"] = bg.batch_num_nodes().numpy(), bg.batch_num_edges().numpy()
    rec["b_pos_enc"], rec["b_fvs"] = bg.ndata["pos_enc"].numpy(), bg.ndata["fvs"].numpy()
    rec["n_graphs"] = np.int64(len(scans))
    np.savez_compressed(os.path.join(out, "graph_pe.npz"), **rec)
    return bg


def _sd_np(model, prefix=""):
    return {prefix + k: v.detach().numpy() for k, v in model.state_dict().items()}


def gen_wiring(bg, out):
    import models as ref_models           # /root/reference/models.py, unmodified
    common = dict(CNN_STUB, out_ch=TINY["out_ch"], fv_dim=TINY["fv_dim"], num_hiddens=TINY["num_hiddens"],
                  node_embed_dim=TINY["node_embed_dim"])
    nets = {
        "gat3": (ref_models.GATNet, dict(common, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                         attn_drop=0.1, negative_slope=0.2)),
        "gat3_nr": (ref_models.GATNet, dict(common, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                            attn_drop=0.1, negative_slope=0.2, res=False)),
        "gat6": (ref_models.GATNet, dict(common, num_gat_layers=6, num_heads=2, num_out_heads=2, feat_drop=0.1,
                                         attn_drop=0.1, negative_slope=0.2, num_hiddens=[16, 8, 8, 8, 8, 8])),
        "gcn3": (ref_models.GCNNet, dict(common, num_gcn_layers=3)),
        "gin3": (ref_models.GINNet, dict(common, num_gin_layers=3)),
        "sage3": (ref_models.SAGENet, dict(common, num_layers=3, feat_drop=0.1, node_ks=[2, 2, 2, 2],
                                           node_sample_rate=0.3)),
        "spgnn3": (ref_models.GATPositionSPGNNNet,
                   dict(common, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1, attn_drop=0.1,
                        negative_slope=0.2, pos_hiddens=TINY["pos_hiddens"], num_pos_heads=1,
                        pos_enc_dim=TINY["pos_enc_dim"])),
        "spgnnnl3": (ref_models.GATPositionSPGNNNet,
                     dict(common, num_gat_layers=3, num_heads=2, num_out_heads=2, feat_drop=0.1, attn_drop=0.1,
                          negative_slope=0.2, pos_hiddens=TINY["pos_hiddens"], num_pos_heads=1,
                          pos_enc_dim=TINY["pos_enc_dim"], mode="PENL")),
    }
    for name, (cls, kw) in nets.items():
        torch.manual_seed(hash(name) % 1000 if False else sum(map(ord, name)))
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            net = cls(**kw)
        # non-zero biases so the bias path is exercised
        with torch.no_grad():
            for k, p in net.named_parameters():
                if k.endswith("bias"):
                    p.normal_(0, 0.1)
                if k.endswith("eps"):
                    p.fill_(0.25)
        net.eval()
        with torch.no_grad():
            res = net(bg)
        rec = {k: v for k, v in _sd_np(net).items()
               if k.split(".")[0] in ("gat", "gcn", "gin", "sage", "gnn_out", "gnn_lobe_out", "gnn_lung_out")}
        rec = {"sd::" + k: v for k, v in rec.items()}
        for j, r in enumerate(res):
            rec[f"out{j}"] = r.numpy()
        # the reference's decision rule on these logits, per tree (job_runner.py:161)
        dec, o = [], 0
        for n in bg.batch_num_nodes().tolist():
            probs = F.softmax(res[0][o:o + n], dim=1)
            max_v, max_idx = torch.max(probs[:, 1:], 0)
            dec.append(max_idx.numpy() + o)
            o += n
        rec["decision"] = np.stack(dec)
        np.savez_compressed(os.path.join(out, f"wiring_{name}.npz"), **rec)
        print(name, [tuple(r.shape) for r in res], sum(v.size for k, v in rec.items() if k.startswith("sd::")), "params")


def main():
    assert os.path.isdir(REF), "run in the authoring container (needs /root/reference)"
    dgl = install_shims()
    sys.path.insert(0, REF)
    R = load_runner_methods(dgl)
    bg = gen_graph_and_pe(R, HERE)
    gen_wiring(bg, HERE)


if __name__ == "__main__":
    main()
