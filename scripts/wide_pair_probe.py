"""Probe of the wide (output layer) GEMM: CTA pairs on/off (set SPGNN_WIDE_PAIR before importing) and the per-head
H = 1 form (N = 256 UMMAs) against the two-head form (N = 128), at the bench size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import wide_check as wc

print("SPGNN_WIDE_PAIR =", os.environ.get("SPGNN_WIDE_PAIR"), torch.cuda.get_device_name(0))
for (M, H, F, k_in, res, act, mode, ng) in [(40001, 2, 1024, 192, 1, 1, 0, 1), (40001, 2, 1024, 192, 1, 1, 1, 2), (38017, 1, 512, 128, 0, 1, 1, 1)]:
    _, e = wc.run(M, H, F, k_in, res, act, mode, n_g=ng)
    print(f"M={M} H={H} F={F} k_in={k_in} res={res} act={act} mode={mode}: err {e[0]:.2e} {e[1]:.2e}", flush=True)
M = 1232896
for H, F in [(2, 1024), (1, 2048)]:
    for mode in (0, 1):
        ms, _ = wc.run(M, H, F, 192, 1, 1, mode, reps=5, check=False)
        print(f"H={H} F={F} mode={mode}: {ms:.2f} ms", flush=True)
