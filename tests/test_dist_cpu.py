"""World-size-2 gloo test (CPU) of the data-parallel-by-graph host logic: shard by graph, per-rank loss normalised
by the GLOBAL class-weight sum, ONE all-reduce of the gradients  ==  the single-process full-batch step."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import TINY_MODELS, golden_graph_inputs, golden_state_dict, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(rec, scans, idx):
    from oracle import dgl_ops
    gs = []
    for i in idx:
        g = dgl_ops.graph_from_adj(scans[i]["adj"])
        g.ndata["fvs"] = torch.from_numpy(scans[i]["fvs"])
        g.ndata["pos_enc"] = torch.from_numpy(rec[f"pos_enc{i}"])
        gs.append(g)
    return dgl_ops.batch(gs), torch.from_numpy(np.concatenate([scans[i]["labels"] for i in idx]))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import models as om
    from spgnn_b200 import dist as sdist
    torch.set_num_threads(1)
    rec, scans = golden_graph_inputs()
    kind, cfg = TINY_MODELS["spgnn3"]
    net = om.GNNNet(kind, cfg)
    net.load_state_dict(golden_state_dict(load_golden("wiring_spgnn3.npz")))
    net.eval()
    lo, hi = sdist.shard_range(len(scans), rank, world)
    bg, y = _batch(rec, scans, range(lo, hi))
    cw = torch.tensor([0.2] + [0.8] * 21)
    logits = net(bg)[0]
    w = cw[y]
    nll = torch.nn.functional.cross_entropy(logits, y, reduction="none")
    local_loss, global_loss = sdist.global_weighted_ce((w * nll).sum(), w.sum())
    local_loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
    sdist.allreduce_sum_(flat)                      # the one gradient collective
    if rank == 0:
        out.put((flat.numpy(), float(global_loss)))
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process():
    from oracle import models as om
    from spgnn_b200 import dist as sdist
    assert [sdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sdist.shard_by_nodes([300, 300, 300, 300], 2) == [0, 2, 4]
    b = sdist.shard_by_nodes([241, 361, 300, 280, 333, 250], 3)
    assert b[0] == 0 and b[-1] == 6 and b == sorted(b)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat2, loss2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    rec, scans = golden_graph_inputs()
    kind, cfg = TINY_MODELS["spgnn3"]
    net = om.GNNNet(kind, cfg)
    net.load_state_dict(golden_state_dict(load_golden("wiring_spgnn3.npz")))
    net.eval()
    bg, y = _batch(rec, scans, range(len(scans)))
    cw = torch.tensor([0.2] + [0.8] * 21)
    loss = torch.nn.functional.cross_entropy(net(bg)[0], y, weight=cw)
    loss.backward()
    flat1 = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None]).numpy()
    assert abs(loss2 - float(loss)) < 1e-5 * abs(float(loss))
    assert np.abs(flat1 - flat2).max() < 1e-5 * np.abs(flat1).max()
