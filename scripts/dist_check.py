"""Multi-GPU check of the data-parallel-by-graph training step on real GPUs (NCCL over NVLink), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_check.py

Every rank first runs the SINGLE-process reference (all trees on its own GPU, no process group), then the ranks
shard the same trees by graph and take the same steps with the one gradient all-reduce and the globally normalised
weighted CE (runner.train_step).  Dropout is off and the CE node mask is explicit, so the two runs are the same
computation up to summation order: parameters after the steps must agree to 1e-5 relative.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spgnn_b200 import dist as sdist, models as sm, ops, pe as spe, runner, synth_device   # noqa: E402

MODEL = dict(out_ch=22, fv_dim=1024, num_hiddens=[256, 128, 64], node_embed_dim=1024, num_gat_layers=3, num_heads=2,
             num_out_heads=2, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, pos_hiddens=[256, 128, 64],
             num_pos_heads=1, pos_enc_dim=39)
TREES, STEPS = 16, 3


def run(first, count, group_ready):
    b = synth_device.make_batch(first, count, ragged=True)
    g = b.graph
    spe.distance_pos_enc(g, pos_enc_dim=39)
    torch.manual_seed(0)
    net = sm.GATPositionSPGNNNet(**MODEL).cuda()
    net.init()
    net.train()
    net.set_gcn_only()
    opt = runner.FlatSGD(net.parameters(), lr=5e-2, momentum=0.9)
    assert opt.world == (dist.get_world_size() if group_ready else 1)
    cw = torch.tensor(runner.CLASS_WEIGHTS_22, device="cuda")
    y = g.ndata["y"]
    # explicit mask: labelled nodes + every 7th node of each tree by LOCAL index (independent of the sharding)
    local = torch.arange(g.num_nodes, device="cuda") - g.node_off[:-1].repeat_interleave(g.batch_num_nodes())
    mask = (y != 0) | (local % 7 == 0)
    losses = []
    for _ in range(STEPS):
        losses.append(float(runner.train_step(net, g, opt, cw, 1.0, mask=mask).item()))
    return opt.flat_p.clone(), losses


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    ref_p, ref_l = run(0, TREES, False)
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    lo, hi = sdist.shard_range(TREES, rank, world)
    p, l = run(lo, hi - lo, True)
    err = float((p.double() - ref_p.double()).abs().max() / ref_p.double().abs().max())
    lerr = max(abs(a - b) / abs(b) for a, b in zip(l, ref_l))
    # every rank must hold the same parameters
    q = p.clone()
    dist.broadcast(q, 0)
    same = bool(torch.equal(q, p))
    print(f"rank {rank}/{world}: trees [{lo},{hi})  loss {l}  ref {ref_l}  param rel err {err:.2e}  loss rel err "
          f"{lerr:.2e}  identical-to-rank0 {same}", flush=True)
    ok = err < 1e-5 and lerr < 1e-5 and same
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if int(t.item()) else "FAILED", flush=True)
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
