"""A few eager 64-tree SPGNN-3 training steps (the reference's TRAIN_BATCH_SIZE) — run under ncu for a launch list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spgnn_b200 import models as sm, ops, pe as spe, runner, synth_device
trees = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = synth_device.make_batch(0, trees, seed=1234, ragged=True).graph
spe.distance_pos_enc(g, pos_enc_dim=39)
model, kind, method, rate = bench.workload(bench.HEADLINE)
torch.manual_seed(0)
net = getattr(sm, method.split(".")[-1])(**model).cuda(); net.init(); net.train(); net.set_gcn_only()
opt = runner.FlatSGD(net.parameters(), lr=5e-4, momentum=0.9)
cw = torch.tensor(runner.CLASS_WEIGHTS_22, device="cuda")
for _ in range(5):
    runner.train_step(net, g, opt, cw, rate)
torch.cuda.synchronize()
