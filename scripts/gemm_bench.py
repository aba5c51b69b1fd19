"""Per-layer timing of the projection GEMMs at the bench shape (N = 4096 trees x 301 nodes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spgnn_b200 import ops
from spgnn_b200._lib import lib, ptr, stream

M = int(os.environ.get("M", 1232896))
LAYERS = [("gat0", 1024, 40, 1028), ("gat1", 512, 256, 516), ("gat2", 256, 128, 260), ("gat_out", 128, 64, 4100),
          ("pgnn0", 40, 0, 514), ("pgnn1", 256, 0, 258), ("pgnn2", 128, 0, 130), ("gnn_out", 1024, 0, 22)]

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

L = lib()
modes = [int(x) for x in os.environ.get("MODES", "1").split(",")]
only = os.environ.get("LAYERS")
if only:
    LAYERS = [l for l in LAYERS if l[0] in only.split(",")]
for name, K1, K2, N in LAYERS:
    x1 = torch.randn(M, K1, device="cuda"); x2 = torch.randn(M, K2, device="cuda") if K2 else None
    K = K1 + K2
    W = torch.randn(N, (K + 3) // 4 * 4, device="cuda")[:, :K]
    y = ops.empty_padded(M, N, "cuda"); dx = ops.empty_padded(M, K1, "cuda"); dW = ops.empty_padded(N, K, "cuda")
    g = torch.randn(M, y.stride(0), device="cuda")[:, :N]
    for mode in modes:
        ws1 = torch.empty(int(L.linear_fwd_ws(N, K1, K2)) + 16, dtype=torch.uint8, device="cuda")
        ws2 = torch.empty(int(L.linear_bwd_input_ws(N, K1)) + 16, dtype=torch.uint8, device="cuda")
        ws3 = torch.empty(int(L.linear_bwd_weight2_ws(M, N, K1, K2)) + 16, dtype=torch.uint8, device="cuda")
        f = lambda: L.linear_fwd(ptr(x1), x1.stride(0), K1, ptr(x2), x2.stride(0) if K2 else 0, K2, ptr(W), W.stride(0), None, 0, 0.0,
                                 ptr(y), y.stride(0), M, N, mode, ptr(ws1), ws1.numel(), stream())
        b = lambda: L.linear_bwd_input(ptr(g), g.stride(0), ptr(W), W.stride(0), 0, ptr(dx), dx.stride(0), M, N, K1, mode, ptr(ws2), ws2.numel(), stream())
        w = lambda: L.linear_bwd_weight2(ptr(g), g.stride(0), ptr(x1), x1.stride(0), K1, ptr(x2), x2.stride(0) if K2 else 0, K2, ptr(dW), dW.stride(0), M, N, ptr(ws3), mode, stream())
        w1 = lambda: L.linear_bwd_weight(ptr(g), g.stride(0), ptr(x1), x1.stride(0), ptr(dW), dW.stride(0), 0, M, N, K1, ptr(ws3), mode, stream())
        tf, tb, tw = timeit(f), timeit(b), timeit(w)
        tw1 = timeit(w1)
        fl = 2.0 * M * N
        print(f"{name:8s} mode {mode} K={K1}+{K2} N={N}: fwd {tf:7.2f} ms {fl*K/tf/1e9:7.1f} TF | dX {tb:7.2f} ms {fl*K1/tb/1e9:7.1f} TF | dW {tw:7.2f} ms {fl*K/tw/1e9:7.1f} TF | dW(v1,src1) {tw1:7.2f} ms", flush=True)
    del x1, x2, y, dx, g
