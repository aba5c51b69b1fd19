"""Per-tensor bisection of the in-degree > 4 gradient deviation.

    python scripts/deg5_bisect.py            # on a GPU box

For a batch of trees with tri-/quadrifurcations (in-degree up to 6) and a bifurcating control of the same kind:

  1. CSR/CSC consistency of the batch (out_slot is a permutation, in_src[out_slot] == owner, out_dst consistent);
  2. ONE GATConv through the planes pipeline, as a hidden layer (flatten) and as an output layer (head mean,
     projection-first = gat_layer.cu, aggregate-first = gat_wide.cu).  The projection output Y and the backward's dY
     planes are captured from the CUDA path; the reference dY is the fp64 autograd gradient of the same layer
     function evaluated ON THE CUDA PATH'S OWN Y.  Errors are printed per column block (dz | G | d el | d er) and
     bucketed by the node's in-degree, so a wrong branch shows up as a bucket;
  3. for the aggregate-first path: d(packed weight) and dX against fp64 autograd, dX bucketed by degree.
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from oracle import dgl_ops, models as om, pe as ope
from spgnn_b200 import graph as sg, models as sm, nn as snn, ops, pe as spe, synth, stack
mods = dict(dgl_ops=dgl_ops, om=om, ope=ope, sg=sg, sm=sm, ops=ops, spe=spe, synth=synth)
F_ = torch.nn.functional


def make_graph(max_children, sizes, seed=7):
    rng = np.random.default_rng(seed)
    return sg.batch_from_adjs([T._random_tree_adj(n, max_children, rng) for n in sizes])


def check_csr(g):
    src, dst = g.src.cpu(), g.dst.cpu()
    in_ptr, in_src, in_eid = g.in_ptr.cpu().long(), g.in_src.cpu().long(), g.in_eid.cpu().long()
    order = torch.sort(dst, stable=True)[1]
    ok = [torch.equal(in_eid, order), torch.equal(in_src, src[order]),
          torch.equal(in_ptr[1:] - in_ptr[:-1], torch.bincount(dst, minlength=g.num_nodes))]
    out_ptr, out_dst, out_slot = g.out_ptr.cpu().long(), g.out_dst.cpu().long(), g.out_slot.cpu().long()
    ok.append(torch.equal(torch.sort(out_slot)[0], torch.arange(g.num_edges)))
    owner = torch.repeat_interleave(torch.arange(g.num_nodes), out_ptr[1:] - out_ptr[:-1])
    ok += [torch.equal(in_src[out_slot], owner), torch.equal(out_dst, dst[in_eid[out_slot]]),
           torch.equal(out_ptr[1:] - out_ptr[:-1], torch.bincount(src, minlength=g.num_nodes))]
    return ok


def ref_layer(Y, src, dst, N, H, F, res, bias, act, slope, mean):
    """GATConv after the projection (DGL 0.7.x) on Y = [z | res | el | er], plain torch."""
    hf = H * F
    z = Y[:, :hf].reshape(N, H, F)
    o = hf
    r = None
    if res:
        r = Y[:, o:o + hf].reshape(N, H, F); o += hf
    el, er = Y[:, o:o + H], Y[:, o + H:o + 2 * H]
    e = F_.leaky_relu(el[src] + er[dst], slope)                             # [E, H]
    m = torch.full((N, H), -float("inf"), dtype=Y.dtype).scatter_reduce(0, dst[:, None].expand(-1, H), e, "amax")
    p = torch.exp(e - m[dst])
    s = torch.zeros(N, H, dtype=Y.dtype).index_add(0, dst, p)
    a = p / s[dst]
    out = torch.zeros(N, H, F, dtype=Y.dtype).index_add(0, dst, a[:, :, None] * z[src])
    if r is not None:
        out = out + r
    if bias is not None:
        out = out + bias.reshape(1, H, F)
    if act == "elu":
        out = F_.elu(out)
    return out.mean(1) if mean else out.reshape(N, hf)


def bucket(err, deg, scale, label):
    """max |err| per node, bucketed by degree, relative to scale"""
    per = err.abs().amax(1) / scale
    parts = []
    for d in sorted(set(deg.tolist())):
        sel = deg == d
        parts.append(f"deg{d}: {float(per[sel].max()):.1e} (n={int(sel.sum())})")
    print(f"      {label:10s} " + "  ".join(parts))


def one_layer(g, D, H, F, act, mean, wide, tag):
    stack.WIDE_OUTPUT_LAYER = wide
    torch.manual_seed(3)
    N = g.num_nodes
    conv = snn.GATConv(D, F, H, 0.0, 0.0, 0.2, True, F_.elu if act == "elu" else None).cuda()
    with torch.no_grad():
        conv.bias.normal_(0, 0.1)
    x = torch.randn(N, D).clamp_(min=0).cuda()
    R = torch.randn(N, F if mean else H * F).cuda()
    plan = stack.StackPlan([stack.LayerPlan(conv, ["x"], "y", mean_heads=mean)], {"x": D}, ["y"])
    assert plan.supported()
    cap = {}
    pl, pbw, wb = stack.planes_linear, stack.planes_linear_bwd_weight, stack._wide_backward

    def cap_linear(A1, W, *a, **k):
        out = pl(A1, W, *a, **k)
        cap.setdefault("Y", out)
        return out

    def cap_bw(dC, X1, X2=None):
        if isinstance(dC, stack.Planes):
            cap.setdefault("dY", dC.float().clone())
        return pbw(dC, X1, X2)

    def cap_wide(*a, **k):
        res = wb(*a, **k)
        cap["wide"] = res
        return res

    stack.planes_linear, stack.planes_linear_bwd_weight, stack._wide_backward = cap_linear, cap_bw, cap_wide
    try:
        _, (out,) = stack.run_stack(plan, g, {"x": x}, False)
        (out * R).sum().backward()
    finally:
        stack.planes_linear, stack.planes_linear_bwd_weight, stack._wide_backward = pl, pbw, wb
    is_wide = "wide" in cap
    src, dst = g.src.cpu(), g.dst.cpu()
    in_ptr = g.in_ptr.cpu().long()
    deg = (in_ptr[1:] - in_ptr[:-1])
    hf = H * F
    print(f"  [{tag}] D={D} H={H} F={F} act={act} mean={mean} wide-path={is_wide}  N={N} max_degree={g.max_degree()}")
    # fp64 autograd reference of the whole layer from x and the parameters
    xd = x.cpu().double().requires_grad_()
    Wfc, Wres = conv.fc.weight.detach().cpu().double().requires_grad_(), conv.res_fc.weight.detach().cpu().double().requires_grad_()
    al, ar = conv.attn_l.detach().cpu().double().requires_grad_(), conv.attn_r.detach().cpu().double().requires_grad_()
    b = conv.bias.detach().cpu().double().requires_grad_()
    z = xd @ Wfc.t()
    el = (z.view(N, H, F) * al).sum(-1)
    er = (z.view(N, H, F) * ar).sum(-1)
    Yr = torch.cat([z, xd @ Wres.t(), el, er], 1)
    outr = ref_layer(Yr, src, dst, N, H, F, True, b, act, 0.2, mean)
    print(f"      forward rel err {float((out.detach().cpu().double() - outr.detach()).abs().max() / outr.abs().max()):.2e}")
    (outr * R.cpu().double()).sum().backward()
    for name, p, r in (("fc.weight", conv.fc.weight, Wfc), ("res_fc.weight", conv.res_fc.weight, Wres),
                       ("attn_l", conv.attn_l, al), ("attn_r", conv.attn_r, ar), ("bias", conv.bias, b)):
        e = float((p.grad.cpu().double() - r.grad).abs().max() / r.grad.abs().max())
        print(f"      grad {name:14s} rel err {e:.2e}")
    if not is_wide:
        Y = cap["Y"].detach().cpu().double()
        Yl = Y[:, :2 * hf + 2 * H].clone().requires_grad_()
        o2 = ref_layer(Yl, src, dst, N, H, F, True, conv.bias.detach().cpu().double(), act, 0.2, mean)
        (o2 * R.cpu().double()).sum().backward()
        dYr, dY = Yl.grad, cap["dY"].cpu().double()[:, :2 * hf + 2 * H]
        for lab, lo, hi in (("dz", 0, hf), ("G(res)", hf, 2 * hf), ("d el", 2 * hf, 2 * hf + H), ("d er", 2 * hf + H, 2 * hf + 2 * H)):
            bucket(dY[:, lo:hi] - dYr[:, lo:hi], deg, float(dYr[:, lo:hi].abs().max()), lab)
    else:
        d_packed, db, dX = cap["wide"]
        dXe = dX.cpu().double()[:, :D] - xd.grad
        bucket(dXe, deg, float(xd.grad.abs().max()), "dX")


def main():
    for name, mc, sizes in (("trifurcations", 4, (150, 90)), ("control", 2, (500, 90)), ("control-small", 2, (150, 90))):
        for tk in ((True, False) if name == "control-small" else (True,)):
            stack.TREE_KERNELS = tk
            g = make_graph(mc, sizes)
            print(f"== {name}: sizes {sizes}, max children {mc}, max_degree {g.max_degree()}, tree kernels {tk}; CSR checks {check_csr(g)}")
            one_layer(g, 128, 2, 64, "elu", False, True, "hidden layer")
            one_layer(g, 256, 2, 128, "elu", False, True, "hidden layer wide rows")
            one_layer(g, 128, 2, 1024, "none", True, False, "output layer, projection-first")
            one_layer(g, 128, 2, 1024, "none", True, True, "output layer, aggregate-first")
            one_layer(g, 192, 2, 1024, "elu", True, True, "SPGNN output layer, aggregate-first")
    stack.TREE_KERNELS = True


if __name__ == "__main__":
    main()
