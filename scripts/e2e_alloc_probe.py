"""Which allocation of the end-to-end loop misses the caching allocator every step?"""
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
torch.cuda.empty_cache()
mode = sys.argv[1] if len(sys.argv) > 1 else "loop"
print("mode", mode)


def stats(tag):
    s = torch.cuda.memory_stats()
    print("%-10s reserved %.2f GB allocated %.2f GB cudaMalloc calls %d frees %d" % (
        tag, torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_allocated() / 1e9, s.get("num_device_alloc", 0),
        s.get("num_device_free", 0)), flush=True)


stats("start")
if mode == "loop":
    i = 0
    for gg in runner.DeviceBatchLoader((hb for _ in range(10)), pos_enc_dim=39, device=dev):
        ls = runner.train_step(net, gg, opt, cw, rate)
        float(ls.item())
        del gg
        stats("step %d" % i); i += 1
elif mode == "upload_only":
    cs = torch.cuda.Stream()
    for i in range(10):
        with torch.cuda.stream(cs):
            bufs = runner._upload(hb, dev)
        for t in bufs: t.record_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        del bufs
        stats("upload %d" % i)
elif mode == "assemble_only":
    bufs = runner._upload(hb, dev); torch.cuda.synchronize()
    for i in range(8):
        gg = runner._assemble(hb, bufs, dev, 39, "dist", defer_checks=True)
        ls = runner.train_step(net, gg, opt, cw, rate)
        float(ls.item())
        del gg
        stats("asm+step %d" % i)
