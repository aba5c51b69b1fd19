"""The oracle's graph arithmetic against LIBRARY sparse kernels it does not share code with: PyTorch's own
``torch.sparse.softmax`` (edge softmax by destination = row softmax of the sparse logit matrix over its specified
entries), ``torch.sparse.mm`` (weighted / normalised neighbourhood sums) and their autograd, plus scipy.sparse for the
GraphConv normalisation — forward AND parameter / input gradients in fp64 (DGL itself is not
installable here, so the arithmetic DGL implements in C++ — SDDMM → edge softmax → SpMM — is pinned against the other
sparse implementation this image has; DGL's own conventions stay pinned by the settings-driven fixtures)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dgl_ops
from spgnn_b200 import synth


def _graph(tree, k, fdim):
    sc = synth.make_scan(tree, k=k, fv_dim=fdim)
    g = dgl_ops.graph_from_adj(sc.adj)
    return g, torch.from_numpy(sc.fvs).double()


def _rows(g, vals):
    """Sparse [N, N] matrix with entry (dst, src) per edge: row v holds v's in-edges."""
    return torch.sparse_coo_tensor(torch.stack([g.dst, g.src]), vals, (g.num_nodes, g.num_nodes)).coalesce()


def gat_sparse(g, x, conv, act):
    n, H, Fo = g.num_nodes, conv._heads, conv._out
    z = (x @ conv.fc.weight.t()).view(n, H, Fo)
    el = (z * conv.attn_l).sum(-1)
    er = (z * conv.attn_r).sum(-1)
    heads = []
    for h in range(H):
        logits = _rows(g, F.leaky_relu(el[g.src, h] + er[g.dst, h], conv.negative_slope))
        att = torch.sparse.softmax(logits, dim=1)                      # library kernel 1
        heads.append(torch.sparse.mm(att, z[:, h, :]))                 # library kernel 2
    out = torch.stack(heads, 1)
    if conv.res_fc is not None:
        out = out + (x @ conv.res_fc.weight.t() if isinstance(conv.res_fc, torch.nn.Linear) else x).view(n, -1, Fo)
    if conv.bias is not None:
        out = out + conv.bias.view(1, H, Fo)
    return act(out) if act is not None else out


@pytest.mark.parametrize("tree,k,heads,out,residual", [(3, 9, 2, 5, True), (5, 40, 1, 7, False), (8, 25, 4, 12, True)])
def test_gatconv_forward_and_gradients_match_torch_sparse(tree, k, heads, out, residual):
    torch.manual_seed(tree)
    g, x = _graph(tree, k, 12)
    conv = dgl_ops.GATConv(12, out, heads, residual=residual, activation=F.elu).double()
    with torch.no_grad():
        conv.bias.normal_()
    xa = x.clone().requires_grad_(True)
    ya = conv(g, xa)
    probe = torch.randn_like(ya)
    ga = torch.autograd.grad((ya * probe).sum(), [xa] + list(conv.parameters()))
    xb = x.clone().requires_grad_(True)
    yb = gat_sparse(g, xb, conv, F.elu)
    gb = torch.autograd.grad((yb * probe).sum(), [xb] + list(conv.parameters()))
    assert torch.allclose(ya, yb, rtol=1e-10, atol=1e-12)
    for a, b, name in zip(ga, gb, ["x"] + [n for n, _ in conv.named_parameters()]):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-11), name


def test_edge_softmax_matches_torch_sparse_softmax_on_a_batch():
    """Several trees batched (dgl.batch): the softmax runs per destination over that node's in-edges only."""
    gs = [_graph(t, k, 4)[0] for t, k in ((1, 6), (2, 30), (3, 11))]
    g = dgl_ops.batch(gs)
    torch.manual_seed(0)
    e = torch.randn(g.number_of_edges(), 3, dtype=torch.float64) * 4
    a = dgl_ops.edge_softmax(g, e)
    for h in range(3):
        want = torch.sparse.softmax(_rows(g, e[:, h]), dim=1).to_dense()
        assert torch.allclose(want[g.dst, g.src], a[:, h], rtol=1e-12, atol=1e-14)
    sums = torch.zeros(g.num_nodes, 3, dtype=torch.float64).index_add_(0, g.dst, a)
    assert torch.allclose(sums, torch.ones_like(sums), atol=1e-12)


def test_graphconv_and_gin_mean_match_scipy_and_torch_sparse():
    import scipy.sparse as sp
    torch.manual_seed(2)
    g, x = _graph(4, 20, 12)
    n = g.num_nodes
    A = sp.csr_matrix((np.ones(g.number_of_edges()), (g.dst.numpy(), g.src.numpy())), shape=(n, n))   # row v: in-edges
    for i, o in ((12, 5), (12, 20)):                       # multiply-first and aggregate-first orders of DGL
        conv = dgl_ops.GraphConv(i, o, activation=None).double()
        with torch.no_grad():
            conv.bias.normal_()
        d_out = np.maximum(np.asarray(A.sum(0)).ravel(), 1.0) ** -0.5     # norm='both': out-degree of the source ...
        d_in = np.maximum(np.asarray(A.sum(1)).ravel(), 1.0) ** -0.5      # ... and in-degree of the destination
        Ahat = sp.diags(d_in) @ A @ sp.diags(d_out)
        want = Ahat @ (x.numpy() @ conv.weight.detach().numpy()) + conv.bias.detach().numpy()
        assert np.allclose(conv(g, x).detach().numpy(), want, rtol=1e-10, atol=1e-12)
    # GINConv 'mean' with eps: (1 + eps) x + mean over in-neighbours, then the MLP (models.py:358-383)
    lin = torch.nn.Linear(12, 6).double()
    gin = dgl_ops.GINConv(lin, "mean", init_eps=0.3, learn_eps=False)
    deg = torch.zeros(n, dtype=torch.float64).index_add_(0, g.dst, torch.ones(g.number_of_edges(), dtype=torch.float64))
    mean_mat = _rows(g, 1.0 / deg[g.dst])
    want = lin(float(1 + gin.eps) * x + torch.sparse.mm(mean_mat, x))      # eps is an fp32 buffer and 1 + eps an fp32 sum, as in DGL
    assert torch.allclose(gin(g, x), want, rtol=1e-10, atol=1e-12)


def test_sage_pool_matches_numpy_maximum_at():
    """SAGEConv('pool'): the max over in-neighbours against numpy's ufunc.at scatter (models.py:668-679)."""
    torch.manual_seed(4)
    g, x = _graph(6, 33, 12)
    n = g.num_nodes
    sage = dgl_ops.SAGEConv(12, 7, "pool", activation=None).double()
    with torch.no_grad():
        sage.bias.normal_()
        m = torch.relu(sage.fc_pool(x)).numpy()
        neigh = np.full((n, 12), -np.inf)
        np.maximum.at(neigh, g.dst.numpy(), m[g.src.numpy()])
        neigh[np.isinf(neigh)] = 0.0                                      # DGL zero-fills nodes without in-edges
        want = x.numpy() @ sage.fc_self.weight.numpy().T + neigh @ sage.fc_neigh.weight.numpy().T + sage.bias.numpy()
        assert np.allclose(sage(g, x).numpy(), want, rtol=1e-10, atol=1e-12)
