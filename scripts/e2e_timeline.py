"""Where does the end-to-end step spend its time?  Phases of runner.DeviceBatchLoader + train_step with CUDA events on
the compute stream and wall clocks on the host: wait-for-copy, decode + graph build, positional encoding, the step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spgnn_b200 import models as sm, ops, pe as spe, runner, synth_device, graph as sg

B = int(os.environ.get("TREES", 4096))
dev = torch.device("cuda", 0)
batch = synth_device.make_batch(0, B, seed=1234, ragged=False)
g = batch.graph
model, kind, method, rate = bench.workload(bench.HEADLINE)
spe.distance_pos_enc(g, pos_enc_dim=39)
torch.manual_seed(0)
net = getattr(sm, method.split(".")[-1])(**model).to(dev); net.init(); net.train(); net.set_gcn_only()
opt = runner.FlatSGD(net.parameters(), lr=5e-4, momentum=0.9)
cw = torch.tensor(runner.CLASS_WEIGHTS_22, device=dev)
for packed in (True, False):
    hb = runner.host_batch_from_graph(g, packed=packed)
    for sync_loss in (True, False):
        for with_copy in (True, False):
            # manual pipeline with phase events
            copy_stream = torch.cuda.Stream()
            def issue():
                cur = torch.cuda.current_stream()
                with torch.cuda.stream(copy_stream):
                    bufs = runner._upload(hb, dev)
                    ev = torch.cuda.Event(); ev.record(copy_stream)
                for t in bufs: t.record_stream(cur)
                return bufs, ev
            steps = 8
            ph = {k: 0.0 for k in ("wait", "assemble", "step")}
            wall = {k: 0.0 for k in ("assemble_cpu", "step_cpu", "loss_sync")}
            pend = issue()
            fixed = runner._upload(hb, dev) if not with_copy else None
            torch.cuda.synchronize()
            t_all = time.perf_counter()
            evs = []
            for i in range(steps):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e[0].record()
                bufs, ev = pend
                torch.cuda.current_stream().wait_event(ev)
                if with_copy and i + 1 < steps:
                    pend = issue()
                e[1].record()
                t0 = time.perf_counter()
                gg = runner._assemble(hb, bufs if with_copy else fixed, dev, 39, "dist")
                e[2].record()
                t1 = time.perf_counter()
                ls = runner.train_step(net, gg, opt, cw, rate)
                e[3].record()
                t2 = time.perf_counter()
                if sync_loss:
                    float(ls.item())
                t3 = time.perf_counter()
                wall["assemble_cpu"] += t1 - t0; wall["step_cpu"] += t2 - t1; wall["loss_sync"] += t3 - t2
                evs.append(e)
            torch.cuda.synchronize()
            total = (time.perf_counter() - t_all) / steps * 1e3
            for e in evs[1:]:
                ph["wait"] += e[0].elapsed_time(e[1]); ph["assemble"] += e[1].elapsed_time(e[2]); ph["step"] += e[2].elapsed_time(e[3])
            n = steps - 1
            print(f"packed={packed} sync_loss={sync_loss} copy_each_step={with_copy}: {total:.1f} ms/step | GPU phases (ms): "
                  + ", ".join(f"{k} {v / n:.1f}" for k, v in ph.items()) + " | host (ms): "
                  + ", ".join(f"{k} {v / steps * 1e3:.1f}" for k, v in wall.items()), flush=True)
    del hb
