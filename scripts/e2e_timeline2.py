"""Host-side cost of every segment of the pipelined end-to-end step, for three places of the next batch's copy
submission: A before this batch's build, B after the build, C after the step's launches."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spgnn_b200 import models as sm, ops, pe as spe, runner, synth_device

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


def issue(chunk=None):
    cur = torch.cuda.current_stream()
    with torch.cuda.stream(copy_stream):
        bufs = runner._upload(hb, dev)
        ev = torch.cuda.Event(); ev.record(copy_stream)
    for t in bufs: t.record_stream(cur)
    return bufs, ev


for order in ("A", "B", "C", "C"):
    steps = 8
    seg = {k: 0.0 for k in ("issue", "assemble", "step", "sync")}
    pend = issue()
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for i in range(steps):
        bufs, ev = pend
        torch.cuda.current_stream().wait_event(ev)
        t0 = time.perf_counter()
        if order == "A" and i + 1 < steps: pend = issue()
        t1 = time.perf_counter()
        gg = runner._assemble(hb, bufs, dev, 39, "dist", defer_checks=True)
        t2 = time.perf_counter()
        if order == "B" and i + 1 < steps: pend = issue()
        t3 = time.perf_counter()
        ls = runner.train_step(net, gg, opt, cw, rate)
        t4 = time.perf_counter()
        if order == "C" and i + 1 < steps: pend = issue()
        t5 = time.perf_counter()
        float(ls.item())
        t6 = time.perf_counter()
        seg["issue"] += (t1 - t0) + (t3 - t2) + (t5 - t4); seg["assemble"] += t2 - t1; seg["step"] += t4 - t3; seg["sync"] += t6 - t5
    torch.cuda.synchronize()
    total = (time.perf_counter() - t_all) / steps * 1e3
    print(f"order {order}: {total:.1f} ms/step | host ms/step: " + ", ".join(f"{k} {v / steps * 1e3:.1f}" for k, v in seg.items()), flush=True)
