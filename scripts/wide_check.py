"""Standalone check + timing of spgnn_wide_linear (aggregate-first output layer GEMM) against fp64.

    python scripts/wide_check.py [--big] [--once]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spgnn_b200._lib import lib, ptr, stream

L = lib()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def planes_of(x, ld):
    M, K = x.shape
    buf = torch.zeros(2, M, ld, dtype=torch.bfloat16, device=x.device)
    L.split_planes(ptr(x), x.stride(0), K, None, 0, 0, 0.0, 0, 0, 0, ptr(buf), ld, M * ld, M, stream())
    return buf


def run(M, H, F, k_in, has_res, act, mode, reps=1, check=True, n_g=1):
    torch.manual_seed(0)
    kp = (k_in + 63) // 64 * 64
    xa = torch.randn(M, (H + 1) * kp, device="cuda")
    for b in range(H + 1):
        xa[:, b * kp + k_in:(b + 1) * kp] = 0
    XA = planes_of(xa, (H + 1) * kp)
    rows = H * F * (1 + has_res)
    W = torch.randn(rows, k_in, device="cuda") / k_in ** 0.5
    bias = torch.randn(H * F, device="cuda") * 0.1
    ws = torch.empty(int(L.wide_linear_ws(H, F, kp, has_res)), dtype=torch.uint8, device="cuda")
    out = torch.empty(M, F, device="cuda")
    outp = torch.zeros(2, M, F, dtype=torch.bfloat16, device="cuda")
    gs = [torch.randn(M, F, device="cuda") for _ in range(n_g)]
    dpre = torch.zeros(2, M, H * F, dtype=torch.bfloat16, device="cuda")
    db = torch.empty(H * F, device="cuda")

    def call():
        gp = [(ptr(g), g.stride(0)) for g in gs] + [(None, 0)] * (3 - n_g)
        L.wide_linear(ptr(XA), XA.shape[2], XA.shape[1] * XA.shape[2], kp, k_in, M, H, F, has_res, ptr(W), W.stride(0),
                      ptr(bias), act, mode, ptr(out), out.stride(0), ptr(outp), F, M * F,
                      gp[0][0], gp[0][1], gp[1][0], gp[1][1], gp[2][0], gp[2][1],
                      ptr(dpre), H * F, M * H * F, ptr(db), ptr(ws), ws.numel(), stream())
    call()
    torch.cuda.synchronize()
    ms = None
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    errs = None
    if check:
        xd = (XA[0].double() + XA[1].double())
        pre = []
        for h in range(H):
            A = torch.cat([xd[:, h * kp:h * kp + k_in]] + ([xd[:, H * kp:H * kp + k_in]] if has_res else []), 1)
            Wh = torch.cat([W[h * F:(h + 1) * F].double()] + ([W[H * F + h * F:H * F + (h + 1) * F].double()] if has_res else []), 1)
            pre.append(A @ Wh.t() + bias[h * F:(h + 1) * F].double())
        f = (lambda t: torch.where(t > 0, t, torch.expm1(t))) if act == 1 else (lambda t: t)
        df = (lambda t: torch.where(t > 0, torch.ones_like(t), torch.exp(t))) if act == 1 else (lambda t: torch.ones_like(t))
        if mode == 0:
            ref = sum(f(p) for p in pre) / H
            errs = (rel(out, ref), rel(outp[0].float() + outp[1].float(), ref))
        else:
            g = sum(x.double() for x in gs) / H
            ref = torch.cat([g * df(p) for p in pre], 1)
            got = dpre[0].double() + dpre[1].double()
            errs = (rel(got, ref), rel(db, ref.sum(0)))
    return ms, errs


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if "--once" in sys.argv:                # for ncu: one launch of each mode at full size
        M = 1232896
        run(M, 2, 1024, 192, 1, 1, 0, check=False)
        run(M, 2, 1024, 192, 1, 1, 1, check=False)
        sys.exit(0)
    worst = 0.0
    for (M, H, F, k_in, res, act, mode, ng) in [(300, 2, 64, 64, 1, 1, 0, 1), (300, 2, 64, 64, 1, 1, 1, 2), (1000, 2, 1024, 192, 1, 1, 0, 1),
                                                (1000, 2, 1024, 192, 1, 1, 1, 1), (777, 2, 1024, 167, 1, 0, 0, 1), (777, 2, 1024, 167, 0, 1, 1, 3),
                                                (513, 1, 512, 128, 0, 1, 0, 1), (513, 4, 96, 128, 1, 1, 1, 1), (4097, 4, 160, 100, 1, 1, 0, 1)]:
        _, e = run(M, H, F, k_in, res, act, mode, n_g=ng)
        worst = max(worst, *e)
        print(f"M={M} H={H} F={F} k_in={k_in} res={res} act={act} mode={mode}: err {e[0]:.2e} {e[1]:.2e}")
    print("worst", worst)
    if "--big" in sys.argv:
        M = 1232896
        for act, mode in [(1, 0), (0, 0), (1, 1), (0, 1)]:
            ms, _ = run(M, 2, 1024, 192, 1, act, mode, reps=5, check=False)
            fl = 2.0 * M * 2048 * 384
            print(f"gat_out wide: act={act} mode={mode}: {ms:.2f} ms ({fl / ms / 1e9:.0f} TF)")
