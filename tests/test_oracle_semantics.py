"""The oracle checked against itself: an independent dense formulation, hand-computed known answers, fp64
gradcheck and structural properties (how to trust the oracle without DGL)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dense, dgl_ops, pe as ope
from spgnn_b200 import synth


def _tree_graph(k=9, tree=3, fdim=12, dtype=torch.float64):
    sc = synth.make_scan(tree, k=k, fv_dim=fdim)
    g = dgl_ops.graph_from_adj(sc.adj)
    adj = torch.zeros(g.num_nodes, g.num_nodes, dtype=dtype)
    adj[g.src, g.dst] = 1
    return g, adj, torch.from_numpy(sc.fvs).to(dtype)


@pytest.mark.parametrize("residual,heads,out", [(True, 2, 5), (False, 1, 7), (True, 1, 12)])
def test_gatconv_matches_dense_masked_softmax(residual, heads, out):
    torch.manual_seed(0)
    g, adj, x = _tree_graph()
    conv = dgl_ops.GATConv(12, out, heads, residual=residual, activation=F.elu).double()
    with torch.no_grad():
        conv.bias.normal_()
    w_res = None
    if residual:
        w_res = conv.res_fc.weight if isinstance(conv.res_fc, torch.nn.Linear) else "identity"
    ref = dense.gat_dense(adj, x, conv.fc.weight, conv.attn_l, conv.attn_r, w_res, conv.bias, 0.2, F.elu, heads, out)
    assert torch.allclose(conv(g, x), ref, rtol=1e-10, atol=1e-12)


def test_gat_known_answer_uniform_attention():
    """3-node path, W = I, attn = 0 ⇒ logits 0 ⇒ uniform attention ⇒ output = mean over the self-looped neighbourhood."""
    adj = np.array([[1, 1, 0], [1, 1, 1], [0, 1, 1]], dtype=np.uint8)
    g = dgl_ops.graph_from_adj(adj)
    conv = dgl_ops.GATConv(2, 2, 1, residual=False, bias=False)
    with torch.no_grad():
        conv.fc.weight.copy_(torch.eye(2))
        conv.attn_l.zero_()
        conv.attn_r.zero_()
    x = torch.tensor([[1.0, 0.0], [0.0, 2.0], [4.0, 4.0]])
    out = conv(g, x).squeeze(1)
    want = torch.stack([(x[0] + x[1]) / 2, (x[0] + x[1] + x[2]) / 3, (x[1] + x[2]) / 2])
    assert torch.allclose(out, want, atol=1e-6)


def test_graphconv_sage_gin_match_dense():
    torch.manual_seed(1)
    g, adj, x = _tree_graph()
    for i, o in ((12, 5), (12, 20)):                       # multiply-first and aggregate-first orders
        conv = dgl_ops.GraphConv(i, o, activation=F.elu).double()
        with torch.no_grad():
            conv.bias.normal_()
        assert torch.allclose(conv(g, x), dense.gcn_dense(adj, x, conv.weight, conv.bias, F.elu), rtol=1e-10, atol=1e-12)
    sage = dgl_ops.SAGEConv(12, 6, "pool", activation=F.elu).double()
    ref = dense.sage_pool_dense(adj, x, sage.fc_pool.weight, sage.fc_pool.bias, sage.fc_self.weight,
                                sage.fc_neigh.weight, sage.bias, F.elu)
    assert torch.allclose(sage(g, x), ref, rtol=1e-10, atol=1e-12)
    gin = dgl_ops.GINConv(None, "mean", init_eps=0.3, learn_eps=True).double()
    assert torch.allclose(gin(g, x), dense.gin_mean_dense(adj, x, gin.eps), rtol=1e-10, atol=1e-12)


def test_gcn_known_answer_star():
    """star with centre 0 and 3 leaves, self loops, W = I, bias 0: centre = Σ x_u / sqrt(4·deg_u)."""
    adj = np.eye(4, dtype=np.uint8)
    adj[0, 1:] = adj[1:, 0] = 1
    g = dgl_ops.graph_from_adj(adj)
    conv = dgl_ops.GraphConv(1, 1)
    with torch.no_grad():
        conv.weight.fill_(1.0)
    x = torch.tensor([[1.0], [2.0], [3.0], [4.0]])
    out = conv(g, x)
    centre = (1 / 2 + (2 + 3 + 4) / np.sqrt(2)) / 2
    assert abs(float(out[0]) - centre) < 1e-6
    assert abs(float(out[1]) - (1 / 2 + 2 / np.sqrt(2)) / np.sqrt(2)) < 1e-6


def test_gatconv_gradcheck_fp64():
    torch.manual_seed(2)
    g, _, _ = _tree_graph(k=4, fdim=5)
    for act in (None, torch.tanh, F.elu):
        for res in (False, True):
            conv = dgl_ops.GATConv(5, 3, 2, residual=res, activation=act).double()
            x = torch.randn(g.num_nodes, 5, dtype=torch.float64, requires_grad=True)
            assert torch.autograd.gradcheck(lambda t: conv(g, t), (x,), eps=1e-6, atol=1e-5)
    e = torch.randn(g.number_of_edges(), 2, 1, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: dgl_ops.edge_softmax(g, t), (e,), eps=1e-6, atol=1e-5)


def test_batch_equals_separate_graphs_and_permutation_equivariance():
    torch.manual_seed(3)
    conv = dgl_ops.GATConv(8, 4, 2, residual=True, activation=F.elu).double()
    parts, outs = [], []
    for t in range(3):
        sc = synth.make_scan(50 + t, k=6 + t, fv_dim=8)
        g = dgl_ops.graph_from_adj(sc.adj)
        g.ndata["fvs"] = torch.from_numpy(sc.fvs).double()
        parts.append(g)
        outs.append(conv(g, g.ndata["fvs"]))
    bg = dgl_ops.batch(parts)
    assert torch.allclose(conv(bg, bg.ndata["fvs"]), torch.cat(outs), atol=1e-12)
    # relabel the nodes of one graph
    g = parts[0]
    perm = torch.randperm(g.num_nodes)
    inv = torch.argsort(perm)
    g2 = dgl_ops.Graph(inv[g.src], inv[g.dst], g.num_nodes)
    out2 = conv(g2, g.ndata["fvs"][perm])
    assert torch.allclose(out2, outs[0][perm], atol=1e-12)


def test_pe_known_answers():
    n = 25                                        # path graph: dist PE = |i - a| / (n - 1)
    adj = np.eye(n, dtype=np.uint8)
    for i in range(n - 1):
        adj[i, i + 1] = adj[i + 1, i] = 1
    pe, _, diam = ope.dist_pos_enc(adj, [0, 7, 24])
    assert diam == n - 1
    want = np.abs(np.arange(n)[:, None] - np.array([0, 7, 24])[None]) / np.float32(n - 1)
    assert np.array_equal(pe, want.astype(np.float32))
    rw = ope.rw_pos_enc(adj, 6)
    assert np.all(rw[:, 0::2] == 0)               # odd powers vanish on a bipartite graph
    assert abs(rw[0, 1] - 0.5) < 1e-7             # 2-step return from an end node: 1 · 1/2
    with pytest.raises(ValueError):
        ope.dist_pos_enc(np.eye(4, dtype=np.uint8), [0])


def test_kink_tape_replays_decisions_and_flags_real_disagreements():
    """oracle/kinks.py: with the oracle's OWN decisions on the tape the result is unchanged; a flipped decision that
    sits on the kink is accepted (and changes the gradient); a flipped decision far from the kink is a violation."""
    from oracle import kinks
    rng = np.random.default_rng(3)
    adj = np.eye(40, dtype=np.uint8)
    for i in range(1, 40):
        p = int(rng.integers(0, i))
        adj[i, p] = adj[p, i] = 1
    g = dgl_ops.graph_from_adj(adj)
    torch.manual_seed(0)
    conv = dgl_ops.GATConv(12, 8, 2, 0.0, 0.0, 0.2, True, F.elu).double()
    x = torch.randn(40, 12, dtype=torch.float64)
    z = (x @ conv.fc.weight.t()).view(40, 2, 8)
    s = ((z * conv.attn_l).sum(-1, keepdim=True)[g.src] + (z * conv.attn_r).sum(-1, keepdim=True)[g.dst]).detach()
    plain = conv(g, x)
    with kinks.use(kinks.Tape([("sign", s > 0)])) as tape:
        same = conv(g, x)
    assert tape.done() and tape.flips == 0 and not tape.violations and torch.equal(plain, same)
    # a decision on the kink: move one logit to +-1e-9 by shifting attn... emulate with a hand-made tape instead
    k = int(s.abs().flatten().argmin())
    flipped = (s > 0).clone()
    flipped.view(-1)[k] = ~flipped.view(-1)[k]
    tol = float(s.abs().flatten()[k] / s.abs().max()) * 1.01
    with kinks.use(kinks.Tape([("sign", flipped)], tol=tol)) as tape:
        conv(g, x)
    assert tape.flips == 1 and not tape.violations
    with kinks.use(kinks.Tape([("sign", flipped)], tol=tol / 2)) as tape:
        conv(g, x)
    assert tape.violations and tape.violations[0][1] == "sign"
    # tape misuse is loud
    with pytest.raises(AssertionError):
        with kinks.use(kinks.Tape([("drop", torch.ones(3))])):
            conv(g, x)
    # dropout replay: the injected mask is what the layer applies
    drop = dgl_ops.GATConv(12, 8, 2, 0.5, 0.0, 0.2, True, None).double()
    drop.train()
    mask = (torch.rand(40, 12, generator=torch.Generator().manual_seed(1)) > 0.5).double() * 2.0
    zd = ((x * mask) @ drop.fc.weight.t()).view(40, 2, 8)
    sd = ((zd * drop.attn_l).sum(-1, keepdim=True)[g.src] + (zd * drop.attn_r).sum(-1, keepdim=True)[g.dst]).detach()
    with kinks.use(kinks.Tape([("drop", mask), ("sign", sd > 0)])) as tape:
        a = drop(g, x)
    drop.eval()
    assert tape.done() and torch.allclose(a, drop(g, x * mask))
    # SAGE max-pool replay: the first-max source per element reproduces the plain result
    sage = dgl_ops.SAGEConv(12, 6, "pool").double()
    m = F.relu(sage.fc_pool(x)).detach()
    arg = torch.full((40, 12), -1, dtype=torch.int64)
    best = torch.full((40, 12), -np.inf, dtype=torch.float64)
    for e in range(g.src.numel()):
        u, v = int(g.src[e]), int(g.dst[e])
        upd = m[u] > best[v]
        best[v] = torch.where(upd, m[u], best[v])
        arg[v] = torch.where(upd, torch.full_like(arg[v], u), arg[v])
    with kinks.use(kinks.Tape([("sign", sage.fc_pool(x).detach() > 0), ("argmax", arg)])) as tape:
        b = sage(g, x)
    assert tape.done() and not tape.violations and torch.allclose(b, sage(g, x))
