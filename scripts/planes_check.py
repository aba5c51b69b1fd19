"""Standalone check + timing of the TMA-fed planes GEMMs against fp64 (run under `timeout`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spgnn_b200._lib import lib, ptr, stream

L = lib()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def to_planes(x, x2=None, p=0.0, seed=0):
    M, K1 = x.shape
    K2 = x2.shape[1] if x2 is not None else 0
    ld = (K1 + K2 + 63) // 64 * 64
    buf = torch.zeros(2, M, ld, dtype=torch.bfloat16, device=x.device)
    L.split_planes(ptr(x), x.stride(0), K1, ptr(x2), x2.stride(0) if x2 is not None else 0, K2, p, seed, 0, 0, ptr(buf), ld,
                   M * ld, M, stream())
    return buf


def planes_float(buf, K):
    return (buf[0].float() + buf[1].float())[:, :K]


def ws(n):
    return torch.empty(int(n), dtype=torch.uint8, device="cuda")


def fwd(A1, K1, A2, K2, W, bias=None, act=0):
    M = A1.shape[1]
    N = W.shape[0]
    C = torch.empty(M, (N + 3) // 4 * 4, device="cuda")[:, :N]
    w = ws(L.planes_linear_fwd_ws(N, K1, K2))
    L.planes_linear_fwd(ptr(A1), A1.shape[2], A1.shape[1] * A1.shape[2], K1, ptr(A2), A2.shape[2] if A2 is not None else 0,
                        A2.shape[1] * A2.shape[2] if A2 is not None else 0, K2, ptr(W), W.stride(0), ptr(bias), act, 0.0,
                        ptr(C), C.stride(0), M, N, ptr(w), w.numel(), stream())
    return C


def bwd_input(dC, N, W, K):
    M = dC.shape[1]
    dA = torch.empty(M, (K + 3) // 4 * 4, device="cuda")[:, :K]
    w = ws(L.planes_linear_bwd_input_ws(N, K))
    L.planes_linear_bwd_input(ptr(dC), dC.shape[2], dC.shape[1] * dC.shape[2], ptr(W), W.stride(0), 0, ptr(dA),
                              dA.stride(0), M, N, K, ptr(w), w.numel(), stream())
    return dA


def bwd_weight(dC, N, X1, K1, X2, K2):
    M = dC.shape[1]
    dW = torch.empty(N, K1 + K2, device="cuda")
    w = ws(L.planes_linear_bwd_weight_ws(M, N, K1, K2))
    L.planes_linear_bwd_weight(ptr(dC), dC.shape[2], dC.shape[1] * dC.shape[2], ptr(X1), X1.shape[2],
                               X1.shape[1] * X1.shape[2], K1, ptr(X2), X2.shape[2] if X2 is not None else 0,
                               X2.shape[1] * X2.shape[2] if X2 is not None else 0, K2, ptr(dW), dW.stride(0), M, N, ptr(w),
                               w.numel(), stream())
    return dW


def check(M, K1, K2, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(M, K1, generator=g).cuda()
    x2 = torch.randn(M, (K2 + 3) // 4 * 4, generator=g).cuda()[:, :K2] if K2 else None
    W = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).cuda()
    go = torch.randn(M, N, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    P1 = to_planes(x1)
    P2 = to_planes(x2) if K2 else None
    Pg = to_planes(go)
    xd = torch.cat([x1, x2], 1).double() if K2 else x1.double()
    e_s = rel(planes_float(P1, K1), x1)
    y = fwd(P1, K1, P2, K2, W, bias, 1)
    torch.cuda.synchronize()
    e_f = rel(y, torch.nn.functional.elu(xd @ W.double().t() + bias.double()))
    dx = bwd_input(Pg, N, W, K1 + K2)
    e_x = rel(dx, go.double() @ W.double())
    dW = bwd_weight(Pg, N, P1, K1, P2, K2)
    torch.cuda.synchronize()
    e_w = rel(dW, go.double().t() @ xd)
    # concatenated single-source form must agree too
    Pc = to_planes(x1, x2) if K2 else None
    e_c = rel(fwd(Pc, K1 + K2, None, 0, W, bias, 1), y) if K2 and K1 % 64 == 0 else 0.0
    print(f"M={M} K={K1}+{K2} N={N}: split {e_s:.1e} fwd {e_f:.2e}  dX {e_x:.2e}  dW {e_w:.2e} cat {e_c:.1e}", flush=True)
    return max(e_f, e_x, e_w)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bench(M, K1, K2, N, name):
    x1 = torch.randn(M, K1, device="cuda")
    x2 = torch.randn(M, (K2 + 3) // 4 * 4, device="cuda")[:, :K2] if K2 else None
    W = torch.randn(N, K1 + K2, device="cuda") / (K1 + K2) ** 0.5
    P1 = to_planes(x1); P2 = to_planes(x2) if K2 else None
    del x1
    go = torch.randn(M, N, device="cuda")
    Pg = to_planes(go)
    del go
    fl = 2.0 * M * N * (K1 + K2) / 1e9
    t_f = timeit(lambda: fwd(P1, K1, P2, K2, W))
    t_x = timeit(lambda: bwd_input(Pg, N, W, K1 + K2))
    t_w = timeit(lambda: bwd_weight(Pg, N, P1, K1, P2, K2))
    print(f"{name}: M={M} K={K1}+{K2} N={N}  fwd {t_f:.2f} ms ({fl / t_f:.0f} TF)  dX {t_x:.2f} ms ({fl / t_x:.0f} TF)  "
          f"dW {t_w:.2f} ms ({fl / t_w:.0f} TF)", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    worst = 0
    for shp in [(128, 64, 0, 16), (300, 64, 0, 48), (1000, 1024, 39, 130), (257, 39, 0, 258), (513, 100, 8, 22),
                (5000, 1024, 40, 1028), (4096, 192, 0, 4100), (3001, 768, 0, 516), (70000, 128, 64, 4100),
                # >= 4 x 74 m-tiles: the 2-CTA cluster (multicast weight tile) path, odd m-tile counts included
                (38001, 512, 256, 516), (40100, 1024, 39, 1028), (37889, 256, 0, 258), (38017, 39, 0, 514)]:
        worst = max(worst, check(*shp))
    print("worst", worst)
    if "--bench" in sys.argv:
        M = 4096 * 301
        bench(M, 1024, 39, 1028, "gat0")
        bench(M, 512, 256, 516, "gat1")
        bench(M, 256, 128, 260, "gat2")
        bench(M, 128, 64, 4100, "gat_out")
        bench(M, 39, 0, 514, "pgnn0")
        bench(M, 256, 0, 258, "pgnn1")
        bench(M, 128, 0, 130, "pgnn2")
        bench(M, 1024, 0, 22, "head")
