"""Device-side breakdown of the end-to-end step: the train step alone, the step with an unrelated H2D copy of the same
size in flight, the decode + build + PE stage alone, and the pipelined loop — CUDA events on the compute stream."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spgnn_b200 import models as sm, pe as spe, runner, synth_device

dev = torch.device("cuda", 0)
g = synth_device.make_batch(0, 4096, seed=1234, ragged=False).graph
model, kind, method, rate = bench.workload(bench.HEADLINE)
spe.distance_pos_enc(g, pos_enc_dim=39)
torch.manual_seed(0)
net = getattr(sm, method.split(".")[-1])(**model).to(dev); net.init(); net.train(); net.set_gcn_only()
opt = runner.FlatSGD(net.parameters(), lr=5e-4, momentum=0.9)
cw = torch.tensor(runner.CLASS_WEIGHTS_22, device=dev)
hb = runner.host_batch_from_graph(g, packed=True)
copy_stream = torch.cuda.Stream()


def timed(fn, n):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def step(): runner.train_step(net, g, opt, cw, rate)
for _ in range(3): step()
print("step alone              %.2f ms" % timed(step, 8), flush=True)


def step_with_copy():
    with torch.cuda.stream(copy_stream):
        bufs = runner._upload(hb, dev)
    step()
    torch.cuda.current_stream().wait_stream(copy_stream)
    del bufs
print("step + concurrent H2D   %.2f ms" % timed(step_with_copy, 8), flush=True)

bufs = runner._upload(hb, dev); torch.cuda.synchronize()
def assemble(): runner._assemble(hb, bufs, dev, 39, "dist", defer_checks=True)
print("decode + build + PE     %.2f ms" % timed(assemble, 8), flush=True)
def assemble_nope(): runner._assemble(hb, bufs, dev, 0, "dist", defer_checks=True)
print("decode + build (no PE)  %.2f ms" % timed(assemble_nope, 8), flush=True)


def both():
    gg = runner._assemble(hb, bufs, dev, 39, "dist", defer_checks=True)
    runner.train_step(net, gg, opt, cw, rate)
print("assemble + step, no copy %.2f ms" % timed(both, 8), flush=True)


def loop(n):
    for gg in runner.DeviceBatchLoader((hb for _ in range(n)), pos_enc_dim=39, device=dev):
        ls = runner.train_step(net, gg, opt, cw, rate)
        float(ls.item())
        del gg
loop(3); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(12); torch.cuda.synchronize()
print("pipelined e2e loop      %.2f ms/step" % ((time.perf_counter() - t0) / 12 * 1e3), flush=True)

# where do the milliseconds between "assemble + step" and the pipelined loop go?  cudaMalloc / cudaFree inside the loop
# (each synchronises the device) and the GPU-idle gap between the end of step i and the first kernel of batch i + 1
st0 = torch.cuda.memory_stats()
gaps, ends = [], []
it = runner.DeviceBatchLoader((hb for _ in range(10)), pos_enc_dim=39, device=dev)
prev_end = None
t0 = time.perf_counter()
host = {"next": 0.0, "step": 0.0, "item": 0.0, "del": 0.0}
while True:
    a = time.perf_counter()
    e_top = torch.cuda.Event(enable_timing=True); e_top.record()
    try:
        gg = next(it)
    except StopIteration:
        break
    b = time.perf_counter()
    ls = runner.train_step(net, gg, opt, cw, rate)
    e_end = torch.cuda.Event(enable_timing=True); e_end.record()
    c = time.perf_counter()
    float(ls.item())
    d = time.perf_counter()
    del gg
    e = time.perf_counter()
    host["next"] += b - a; host["step"] += c - b; host["item"] += d - c; host["del"] += e - d
    ends.append((e_top, e_end))
torch.cuda.synchronize()
n = len(ends)
print("loop again: %.2f ms/step; host ms/step: %s" % ((time.perf_counter() - t0) / n * 1e3,
      ", ".join("%s %.2f" % (k, v / n * 1e3) for k, v in host.items())))
print("GPU time from loop top to end of step (per step): " + " ".join("%.1f" % a.elapsed_time(b) for a, b in ends))
st1 = torch.cuda.memory_stats()
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_sync_all_streams"):
    print(k, st1.get(k, 0) - st0.get(k, 0))
print("reserved GB", torch.cuda.memory_reserved() / 1e9, "allocated GB", torch.cuda.memory_allocated() / 1e9)
