"""Per-kernel-family time of one bench step (CUDA events around every C-ABI call), for quick A/B runs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spgnn_b200 import models as sm, ops, pe as spe, runner, synth_device
from spgnn_b200._lib import lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
batch = synth_device.make_batch(first_tree=0, count=B, seed=bench.SEED)
g = batch.graph
spe.distance_pos_enc(g, pos_enc_dim=39)
torch.manual_seed(0)
net = sm.GATPositionSPGNNNet(**bench.MODEL).to(dev)
net.init(); net.train(); net.set_gcn_only()
opt = runner.FlatSGD(net.parameters(), lr=bench.LR, momentum=bench.MOMENTUM)
cw = torch.tensor(runner.CLASS_WEIGHTS_22, device=dev)
ops.manual_seed(1)
for _ in range(3):
    runner.train_step(net, g, opt, cw, 0.15)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    runner.train_step(net, g, opt, cw, 0.15)
e1.record(); torch.cuda.synchronize()
L = lib(); L.profile = []
for _ in range(3):
    runner.train_step(net, g, opt, cw, 0.15)
torch.cuda.synchronize()
prof, L.profile = L.profile, None
agg = {}
for name, key, a, b in prof:
    agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / 3
print(f"step {e0.elapsed_time(e1)/5:.2f} ms | " + " ".join(f"{k}={v:.2f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
