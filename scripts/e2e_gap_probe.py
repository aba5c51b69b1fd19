"""Where does the pipelined loop lose time against (decode + build + PE + step)?  Per-step host timestamps and GPU events."""
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
del g
variant = sys.argv[1] if len(sys.argv) > 1 else "loader"


def run(n, log):
    it = runner.DeviceBatchLoader((hb for _ in range(n)), pos_enc_dim=39, device=dev)
    rows = []
    while True:
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        try:
            gg = next(it)
        except StopIteration:
            break
        t1 = time.perf_counter()
        e1 = torch.cuda.Event(enable_timing=True); e1.record()          # after decode + build + PE (+ next copy issued)
        ls = runner.train_step(net, gg, opt, cw, rate)
        e2 = torch.cuda.Event(enable_timing=True); e2.record()
        t2 = time.perf_counter()
        e2.synchronize()
        t3 = time.perf_counter()
        float(ls.item())
        t4 = time.perf_counter()
        del gg
        rows.append((t0, t1, t2, t3, t4, e0, e1, e2))
    if log:
        for i, (t0, t1, t2, t3, t4, e0, e1, e2) in enumerate(rows):
            nxt = rows[i + 1][0] if i + 1 < len(rows) else t4
            print("step %d: host next %.2f launch %.2f wait %.2f item %.3f | GPU assemble %.2f step %.2f | wall %.2f" % (
                i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, e0.elapsed_time(e1), e1.elapsed_time(e2),
                (nxt - t0) * 1e3), flush=True)


run(3, False)
torch.cuda.synchronize()
run(8, True)
