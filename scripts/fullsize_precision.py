"""Where does the full-size gradient differ from the sum over sub-batches (tests/test_gpu_fullsize.py)?

1. spgnn_planes_linear_bwd_weight (dW = dC^T X, reduction over M node rows) against an fp64 matmul on the device at
   M = 77 k ... 1.23 M rows, random operands with the statistics of the step (dC zero-mean, X = relu(normal)):
   error relative to the largest entry, and the same for the sum over 16 row blocks;
2. decisions (LeakyReLU branches) of the SPGNN-3 forward on a 4096-tree batch vs the same trees in sub-batches.
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from spgnn_b200 import models as sm, ops, pe as spe, stack, synth_device
from helpers import FULL_MODELS

torch.manual_seed(0)
dev = "cuda"


def tn_case(M, N, K, signed_x=False):
    gen = torch.Generator(device=dev).manual_seed(M + N)
    dC = torch.randn(M, N, device=dev, generator=gen)
    X = torch.randn(M, K, device=dev, generator=gen)
    if not signed_x:
        X.clamp_(min=0)
    dCp, Xp = stack.split_planes(dC), stack.split_planes(X)
    dW = stack.planes_linear_bwd_weight(dCp, Xp).double()
    # reference from the planes' own values (what the GEMM is given), fp64
    ref = dCp.float().double().t() @ Xp.float().double()
    e_big = float((dW - ref).abs().max() / ref.abs().max())
    acc = torch.zeros_like(ref)
    step = (M + 15) // 16
    for o in range(0, M, step):
        a = stack.split_planes(dC[o:o + step]); b = stack.split_planes(X[o:o + step])
        acc += stack.planes_linear_bwd_weight(a, b).double()
    e_sum = float((acc - ref).abs().max() / ref.abs().max())
    e_bs = float((acc - dW).abs().max() / ref.abs().max())
    print(f"TN M={M:8d} N={N:4d} K={K:4d} signed_x={signed_x}: big vs fp64 {e_big:.2e}   sum of 16 blocks vs fp64 {e_sum:.2e}   big vs sum {e_bs:.2e}", flush=True)


for M in (77056, 308224, 1232896):
    tn_case(M, 516, 1063)
    tn_case(M, 516, 1063, signed_x=True)
tn_case(1232896, 132, 384)
tn_case(1232896, 22, 1024)

# 2. decisions big vs sub-batches
kind, cfg = FULL_MODELS["st_pgat_spgnn_3"]
B = 4096
big = synth_device.make_batch(0, B, ragged=True)
spe.distance_pos_enc(big.graph, pos_enc_dim=39)
net = sm.GATPositionSPGNNNet(**cfg).cuda(); net.init(); net.eval()
ops.KINK_TRACE = []
with torch.no_grad():
    out_big = net(big.graph)
rec_big, ops.KINK_TRACE = ops.KINK_TRACE, None
eoff = big.graph.edge_off.cpu().numpy(); noff = big.graph.node_off.cpu().numpy()
flips = [0] * len(rec_big); total = [0] * len(rec_big); maxdiff = 0.0
for first in range(0, B, 512):
    sb = synth_device.make_batch(first, 512, ragged=True)
    spe.distance_pos_enc(sb.graph, pos_enc_dim=39)
    ops.KINK_TRACE = []
    with torch.no_grad():
        out = net(sb.graph)
    rec, ops.KINK_TRACE = ops.KINK_TRACE, None
    for i, ((k, a), (_, b)) in enumerate(zip(rec, rec_big)):
        bb = b[eoff[first]:eoff[first + 512]]
        flips[i] += int((a != bb).sum()); total[i] += a.numel()
    for a, b in zip(out, out_big):
        maxdiff = max(maxdiff, float((a - b[noff[first]:noff[first + 512]]).abs().max()))
print("decision records:", len(rec_big), "flips per layer:", flips, "of", total, " max |out_sub - out_big| =", maxdiff)
