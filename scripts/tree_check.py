"""Per-tree layer kernels (gat_tree.cu) against the chunk kernels (gat_layer.cu) on the same inputs.

    python scripts/tree_check.py fwd|bwd [B] [H] [F] [res] [ng]
"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spgnn_b200 import stack, synth_device, ops
from spgnn_b200._lib import lib, ptr, stream

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2
F = int(sys.argv[4]) if len(sys.argv) > 4 else 64
res = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ng = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ragged = int(sys.argv[7]) if len(sys.argv) > 7 else 1
dev = torch.device("cuda", 0)
g = synth_device.make_batch(first_tree=0, count=B, seed=1, ragged=bool(ragged)).graph
N, E = g.num_nodes, g.num_edges
print("N", N, "E", E, "max_nodes", g.max_nodes, "max_degree", g.max_degree(), flush=True)
HF = H * F
ycols = HF * (1 + res) + 2 * H
torch.manual_seed(0)
Y = torch.empty(N, (ycols + 31) // 32 * 32, device=dev)[:, :ycols]
Y.copy_(torch.randn(N, ycols, device=dev))
bias = torch.randn(HF, device=dev) * 0.1
L_ = lib()
stack._check_abi()


def desc(tree):
    d = stack._Layer()
    d.in_ptr, d.in_src = ptr(g.in_ptr), ptr(g.in_src)
    d.out_ptr, d.out_dst, d.out_slot = ptr(g.out_ptr), ptr(g.out_dst), ptr(g.out_slot)
    d.N, d.H, d.F = N, H, F
    if tree:
        d.node_off, d.B, d.max_nodes, d.max_degree = ptr(g.node_off), g.batch_size, g.max_nodes, g.max_degree()
    d.Y, d.ldy, d.res_off, d.el_off, d.er_off = ptr(Y), Y.stride(0), HF, HF * (1 + res), HF * (1 + res) + H
    d.res_mode, d.act, d.negative_slope, d.mean_heads = res, 1, 0.2, 0
    d.bias = ptr(bias)
    d.attn_drop_p, d.attn_seed = 0.1, 77
    return d


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


outs = {}
for tree in (0, 1):
    att = torch.zeros(E, H, device=dev)
    out32 = ops.empty_padded(N, HF, dev)
    P = stack.Planes(N, HF, dev)
    d = desc(tree)
    d.att = ptr(att)
    d.out, d.ldo = ptr(out32), out32.stride(0)
    d.n_sinks = 1
    sk = d.sinks[0]
    sk.hi, sk.ld, sk.plane_stride, sk.concat_chunks, sk.chunk_off, sk.drop_p, sk.seed = P.ptr(), P.ld, P.ps, (HF + 3) // 4 + 5, 5, 0.1, 123
    print("fwd tree =", tree, flush=True)
    L_.gat_layer_fwd(ctypes.byref(d), stream())
    torch.cuda.synchronize()
    res_f = [att, out32.clone(), P.float()]
    if what == "bwd":
        gs = [torch.randn(N, HF + 8, device=dev) for _ in range(ng)]
        torch.manual_seed(5)
        for x in gs:
            x.copy_(torch.randn(N, HF + 8, device=dev))
        dY = stack.Planes(N, ycols, dev)
        dY.buf.zero_()
        d.n_gsrc = ng
        for s_, x in enumerate(gs):
            k = d.gsrc[s_]
            k.g, k.ld, k.concat_chunks, k.chunk_off, k.drop_p, k.seed = x.data_ptr() + 4 * 8, x.stride(0), (HF + 8) // 4, 2, 0.1 * s_, 55 + s_
        d.dY_hi, d.dY_ld, d.dY_ps = dY.ptr(), dY.ld, dY.ps
        g_ws = torch.empty(N, HF, device=dev)
        ds = torch.empty(E * H, device=dev)
        db = torch.empty(HF, device=dev)
        dbw = torch.empty(int(L_.gat_layer_dbias_ws(N, H, F)), dtype=torch.uint8, device=dev)
        d.g_ws, d.ds_ws, d.dbias, d.dbias_ws = ptr(g_ws), ptr(ds), ptr(db), ptr(dbw)
        print("bwd tree =", tree, flush=True)
        L_.gat_layer_bwd(ctypes.byref(d), stream())
        torch.cuda.synchronize()
        res_f += [dY.float(), db]
    outs[tree] = res_f
names = ["att", "out", "planes", "dY", "dbias"]
for nme, a, b in zip(names, outs[1], outs[0]):
    print(f"{nme}: rel err tree vs chunk {rel(a, b):.2e}")
