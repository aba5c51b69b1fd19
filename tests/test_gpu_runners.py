"""exp_settings-driven runners on the GPU: a short synthetic training run, checkpoint round trip, inference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _settings(tmp_path, preset="st_pgat_spgnn_3", scans=6):
    from spgnn_b200.settings import Settings
    s = Settings(preset)
    s.SYNTHETIC_SCANS, s.SYNTHETIC_RAGGED = scans, True
    s.MODEL_ROOT_PATH, s.EXP_NAME = str(tmp_path), "exp"
    s.TRAIN_BATCH_SIZE, s.TRAIN_SAMPLE_SIZE, s.GCN_STEPS, s.LOG_STEPS, s.NUM_EPOCHS, s.SAVE_EPOCHS = 4, 4, 12, 1, 2, 1
    s.OPTIMIZER = dict(s.OPTIMIZER, lr=5e-3)
    return s


def test_train_checkpoint_test_cycle(tmp_path):
    from spgnn_b200 import job_runner, ops
    from spgnn_b200.settings import get_callable_by_name
    torch.manual_seed(0)
    ops.manual_seed(0)
    s = _settings(tmp_path)
    tr = get_callable_by_name(s.JOB_RUNNER_CLS)(s)
    assert isinstance(tr, job_runner.GCNTrainSPGNN)
    assert all(p.requires_grad for p in tr.model.parameters())
    hist = tr.run()
    assert len(hist) == 2 * 12 and np.isfinite(hist).all()
    assert np.mean(hist[-4:]) < np.mean(hist[:4])                  # the loss goes down on a fixed batch
    ck = sorted((tmp_path / "exp").glob("*.pth"))
    assert ck, "no checkpoint written"
    state = torch.load(ck[-1], weights_only=False)
    assert {"iteration", "epoch_n", "model_dict", "metric"} <= set(state)
    assert "gat.gat_layers.0.fc.weight" in state["model_dict"]     # DGL key names

    s2 = _settings(tmp_path)
    s2.RELOAD_CHECKPOINT_PATH = str(ck[-1])
    te = get_callable_by_name(s2.TEST_RUNNER_CLS)(settings_module=s2, output_path=str(tmp_path / "out"))
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v, te.model.state_dict()[k]), k
    res = te.run()
    assert len(res) == 6 and (tmp_path / "out" / "gnn_predictions.pkl").exists()
    r = next(iter(res.values()))
    assert r["decision"].shape == (21,) and 0.0 <= r["acc"] <= 1.0


@pytest.mark.parametrize("preset", ["st_gat_3", "st_gcn_3", "st_gin_3", "st_sage_3", "st_gat_6_nr"])
def test_other_presets_take_a_training_step(tmp_path, preset):
    from spgnn_b200.settings import get_callable_by_name
    s = _settings(tmp_path, preset, scans=3)
    tr = get_callable_by_name(s.JOB_RUNNER_CLS)(s)
    from spgnn_b200 import job_runner
    assert isinstance(tr, job_runner.GCNTrainSAGE) == (preset == "st_sage_3")     # neighbour-sampled mini-batches
    losses = tr.train_batch([tr.source(u) for u in tr.source.uids], steps=3)
    assert len(losses) == 3 and np.isfinite(losses).all()
    assert np.isfinite(tr.validate(tr.source.uids[:1]))


def test_stage1_pickles_in_the_reference_layout_feed_the_device_batch(tmp_path):
    """ConvEmbeddingExtractor's pickle (job_runner.py:796-805: ``fvs`` float64 [n,1024], ``adj`` uint8 [n,n],
    ``labels`` uint8 [n], ``fvs_out`` float64 [n,22] + ref / all_airway / branch_info / meta) written to
    ``<DB_PATH>/derived/conv_embedding/<uid>.pkl`` and read back through ScanSource → host_batch → the device batch
    builder: graph, features and labels equal those built straight from the arrays."""
    import pickle
    from spgnn_b200 import graph as sg, job_runner, runner, synth
    from spgnn_b200.settings import Settings
    scans = synth.make_scans(50, 3, ragged=True)
    root = tmp_path / "derived" / "conv_embedding"
    root.mkdir(parents=True)
    for i, sc in enumerate(scans):
        n = sc.adj.shape[0]
        with open(root / f"scan_{i:03d}.pkl", "wb") as fp:
            pickle.dump({"fvs": sc.fvs.astype(np.float64), "adj": sc.adj.astype(np.uint8),
                         "labels": sc.labels.astype(np.uint8), "fvs_out": sc.fvs_out.astype(np.float64),
                         "ref": np.zeros((4, 4, 4), np.uint8), "all_airway": np.zeros((4, 4, 4), np.uint8),
                         "branch_info": {k: dict(label=int(sc.labels[k])) for k in range(n)},
                         "meta": {"uid": f"scan_{i:03d}", "spacing": [1.0, 1.0, 1.0]}}, fp)
    s = Settings("st_pgat_spgnn_3")
    s.DB_PATH = str(tmp_path)
    src = job_runner.ScanSource(s)
    assert src.uids == [f"scan_{i:03d}" for i in range(3)]
    loaded = [src(u) for u in src.uids]
    assert loaded[0]["fvs"].dtype == np.float64 and loaded[0]["labels"].dtype == np.uint8
    hb = job_runner.host_batch(loaded, packed=False)
    assert hb.fvs.dtype == torch.float32 and hb.labels.dtype == torch.int64 and hb.adj_cat.dtype == torch.uint8
    g = runner.batch_to_device(job_runner.host_batch(loaded), pos_enc_dim=39)        # the loader's packed wire format
    ref = sg.batch_from_adjs([sc.adj for sc in scans])
    assert torch.equal(g.src, ref.src) and torch.equal(g.dst, ref.dst) and torch.equal(g.in_src, ref.in_src)
    assert torch.equal(g.ndata["fvs"].cpu(), torch.from_numpy(np.concatenate([sc.fvs for sc in scans])))
    assert torch.equal(g.ndata["y"].cpu(), torch.from_numpy(np.concatenate([sc.labels for sc in scans]).astype(np.int64)))
    assert g.ndata["pos_enc"].shape == (g.num_nodes, 39)


def test_resume_restores_momentum_lr_and_iteration(tmp_path):
    """save_model / reload_model_from_cache (job_runner.py:298-350): optimizer_dict and scheduler_dict are in the
    checkpoint; a run resumed with them in RELOAD_DICT_LIST continues with the same momentum buffer, learning rate
    and iteration + 1; without RELOAD_CHECKPOINT nothing is loaded; a missing checkpoint is not an error."""
    from spgnn_b200 import job_runner, ops
    from spgnn_b200.settings import get_callable_by_name
    torch.manual_seed(0)
    ops.manual_seed(0)
    s = _settings(tmp_path, "st_gat_3", scans=3)
    s.RELOAD_CHECKPOINT = True                      # nothing to reload yet: logged, training starts from scratch
    tr = get_callable_by_name(s.JOB_RUNNER_CLS)(s)
    assert tr.current_iteration == 0
    tr.train_batch([tr.source(u) for u in tr.source.uids], steps=3)
    tr.optimizer.set_lr(tr.optimizer.lr * tr.gamma)
    tr.epoch_n = 1
    path = tr.save_model(metric=0.5)
    state = torch.load(path, weights_only=False)
    assert {"iteration", "epoch_n", "model_dict", "optimizer_dict", "scheduler_dict", "metric"} <= set(state)
    assert len(state["optimizer_dict"]["state"]) == len(tr.optimizer.params)

    s2 = _settings(tmp_path, "st_gat_3", scans=3)
    s2.RELOAD_CHECKPOINT = True
    s2.RELOAD_DICT_LIST = ["model_dict", "metric", "optimizer_dict", "scheduler_dict"]
    tr2 = get_callable_by_name(s2.JOB_RUNNER_CLS)(s2)
    assert tr2.current_iteration == tr.current_iteration + 1 and tr2.epoch_n == 1
    assert abs(tr2.optimizer.lr - tr.optimizer.lr) < 1e-12 and tr2.metric == {"metric": 0.5}
    assert torch.equal(tr2.optimizer.buf, tr.optimizer.buf) and torch.equal(tr2.optimizer.flat_p, tr.optimizer.flat_p)
    assert all(tr2.optimizer.has_buf)

    s3 = _settings(tmp_path, "st_gat_3", scans=3)   # RELOAD_CHECKPOINT False: a path alone does not reload
    s3.RELOAD_CHECKPOINT_PATH = path
    tr3 = get_callable_by_name(s3.JOB_RUNNER_CLS)(s3)
    assert tr3.current_iteration == 0 and not torch.equal(tr3.optimizer.flat_p, tr.optimizer.flat_p)
    with pytest.raises(job_runner.SpgnnError):
        s4 = _settings(tmp_path, "st_gat_3", scans=3)
        s4.OPTIMIZER = dict(s4.OPTIMIZER, groups={})
        get_callable_by_name(s4.JOB_RUNNER_CLS)(s4)


def test_packed_wire_format_decodes_to_the_dense_batch_bit_for_bit():
    """runner.HostBatch(packed=True) (csrc/wire.cu: zero-suppressed fvs, int32 edge lists, uint8 labels) → device
    batch: graph arrays, features (including -0.0 and denormals), labels and positional encoding are bit-identical
    to the dense stage-1 layout, at about half the bytes; trees with up to 6 children and a 1-node tree included."""
    import test_gpu_parity as T
    from spgnn_b200 import job_runner, runner
    rng = np.random.default_rng(5)
    adjs = [T._random_tree_adj(n, mc, rng) for n, mc in ((301, 2), (25, 6), (1, 2), (120, 3), (64, 2))]
    scans = []
    for a in adjs:
        n = a.shape[0]
        f = np.maximum(rng.standard_normal((n, 1024)), 0).astype(np.float32)
        f[0, :4] = [-0.0, 1e-42, np.float32(3.5), 0.0]                       # negative zero and a denormal survive
        scans.append(dict(adj=a, fvs=f, fvs_out=rng.standard_normal((n, 22)).astype(np.float32),
                          labels=rng.integers(0, 22, n).astype(np.uint8)))
    dense = job_runner.host_batch(scans, packed=False)
    packed = job_runner.host_batch(scans, packed=True)
    assert packed.nbytes() < 0.6 * dense.nbytes()
    gd = runner.batch_to_device(dense, pos_enc_dim=0)
    gp = runner.batch_to_device(packed, pos_enc_dim=0)
    for k in ("src", "dst", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_slot", "node_off", "edge_off"):
        assert torch.equal(getattr(gd, k), getattr(gp, k)), k
    assert gp.max_degree() == gd.max_degree() > 4
    assert torch.equal(gd.ndata["fvs"].view(torch.int32), gp.ndata["fvs"].view(torch.int32))       # bit patterns
    assert torch.equal(gd.ndata["fvs_out"], gp.ndata["fvs_out"]) and torch.equal(gd.ndata["y"], gp.ndata["y"])
    assert gp.ndata["y"].dtype == torch.int64
    # an all-zero feature matrix and a dense one
    for f in (np.zeros((40, 96), np.float32), rng.standard_normal((40, 96)).astype(np.float32)):
        sc = [dict(adj=T._random_tree_adj(40, 2, rng), fvs=f, fvs_out=np.zeros((40, 22), np.float32),
                   labels=np.zeros(40, np.int64))]
        g1 = runner.batch_to_device(job_runner.host_batch(sc, packed=True), pos_enc_dim=0)
        assert torch.equal(g1.ndata["fvs"].cpu(), torch.from_numpy(f))
    # through the pipelined loader, with the positional encoding
    big = [s for s in scans if s["adj"].shape[0] >= 21]
    a = next(iter(runner.DeviceBatchLoader([job_runner.host_batch(big, packed=True)], pos_enc_dim=39)))
    b = next(iter(runner.DeviceBatchLoader([job_runner.host_batch(big, packed=False)], pos_enc_dim=39)))
    assert torch.equal(a.ndata["pos_enc"], b.ndata["pos_enc"]) and torch.equal(a.ndata["fvs"], b.ndata["fvs"])


def test_the_reference_training_loop_runs_verbatim_on_the_models():
    """The body of GCNTrainSPGNN.train's GCN_STEPS loop (job_runner.py:1885-1919) written as the reference writes it —
    numpy sampling table, Python list-comprehension mask, ``F.cross_entropy(gnn_out[mask], labels[mask], weight=)``,
    ``torch.optim.SGD`` — on a spgnn_b200 model, against runner.train_step (fused masked CE + FlatSGD) fed the same
    masks: same losses step by step and the same parameters afterwards."""
    import torch.nn.functional as F
    import test_gpu_parity as T
    from spgnn_b200 import graph as sg, models as sm, ops, pe as spe, runner, synth
    from spgnn_b200.settings import PRESETS
    cfg = dict(PRESETS["st_pgat_spgnn_3"]["MODEL"])
    cfg.pop("method")
    cfg.update(feat_drop=0.0, attn_drop=0.0)
    m = dict(sg=sg, synth=synth)
    scans = T._scan_dicts(m, 40, 5)
    batch_g = T._device_batch(m, scans)
    batch_g.ndata["y"] = torch.from_numpy(np.concatenate([s["labels"] for s in scans]).astype(np.int64)).cuda()
    spe.distance_pos_enc(batch_g, pos_enc_dim=39)
    torch.manual_seed(0)
    model = sm.GATPositionSPGNNNet(**cfg).cuda()
    model.init(); model.train(); model.set_gcn_only()
    twin = sm.GATPositionSPGNNNet(**cfg).cuda()
    twin.load_state_dict(model.state_dict()); twin.train(); twin.set_gcn_only()
    GCN_STEPS, sampling_rate = 4, 0.3
    weight_t = torch.tensor(runner.CLASS_WEIGHTS_22, device="cuda")
    optimizer = torch.optim.SGD(model.parameters(), lr=5e-3, momentum=0.9)
    flat = runner.FlatSGD(twin.parameters(), lr=5e-3, momentum=0.9)
    # ---- reference lines
    n_nodes = [batch_g.number_of_nodes()]
    labels_list_cat = batch_g.ndata['y']
    sampling_t = torch.ones_like(labels_list_cat, dtype=torch.float32) * sampling_rate
    sampling_t[labels_list_cat.nonzero(as_tuple=True)] = 1.0
    sampling_t = sampling_t.detach().cpu().numpy()
    np.random.seed(3)
    random_list = np.random.random_sample(GCN_STEPS * sum(n_nodes)).reshape(GCN_STEPS, sum(n_nodes))
    ref_losses, our_losses = [], []
    for n in range(GCN_STEPS):
        optimizer.zero_grad()
        mask = [rn < sampling_t[k] for k, rn in zip(range(sum(n_nodes)), random_list[n])]
        assert (all([mask[x.item()] for x in labels_list_cat.nonzero()]))
        gnn_out, n_embed, p_embed = model(batch_g)
        loss_gnn = F.cross_entropy(gnn_out[mask], labels_list_cat[mask], weight=weight_t)
        loss_gnn.backward()
        ref_losses.append(loss_gnn.item())
        optimizer.step()
        # ---- the runner's step on the twin, same mask
        mt = torch.from_numpy(np.asarray(mask)).cuda()
        our_losses.append(float(runner.train_step(twin, batch_g, flat, weight_t, sampling_rate, mask=mt).item()))
    assert np.isfinite(ref_losses).all() and ref_losses[-1] < ref_losses[0]
    assert np.allclose(ref_losses, our_losses, rtol=2e-5, atol=1e-6), (ref_losses, our_losses)
    for (k, a), b in zip(model.state_dict().items(), twin.state_dict().values()):
        assert T.rel_err(b, a) < 1e-4, k


def test_seed_salt_changes_every_mask_and_resets():
    """spgnn_seed_salt_set: the same (p, seed) draws a different dropout mask under a different salt (what lets a
    CUDA-graph replay of a step draw fresh masks although its seeds are frozen kernel arguments), salt 0 restores it."""
    from spgnn_b200 import ops
    from spgnn_b200._lib import lib, ptr, stream
    assert lib()._dll.spgnn_seed_salt_units() >= 8
    salt = torch.zeros(1, dtype=torch.int64, device="cuda")
    base = ops.drop_mask(500, 64, 0.3, 1234, "cuda")
    try:
        salt.fill_(5)
        lib().seed_salt_set(ptr(salt), stream())
        other = ops.drop_mask(500, 64, 0.3, 1234, "cuda")
        assert not torch.equal(base, other) and abs(float((other > 0).float().mean()) - 0.7) < 0.02
        again = ops.drop_mask(500, 64, 0.3, 1234, "cuda")
        assert torch.equal(other, again)
    finally:
        salt.zero_()
        lib().seed_salt_set(ptr(salt), stream())
    assert torch.equal(ops.drop_mask(500, 64, 0.3, 1234, "cuda"), base)


def test_cuda_graph_step_matches_eager_and_redraws_masks():
    """runner.GraphedTrainStep on a 6-tree batch: (a) without any randomness (dropout 0, every node kept) the replayed
    steps reproduce the eager steps; (b) with dropout two objects built from the same state replay the same sequence
    (deterministic; that the salt redraws every mask is test_seed_salt_changes_every_mask_and_resets)."""
    from spgnn_b200 import models as sm, ops, pe as spe, runner, synth_device
    from helpers import FULL_MODELS, rel_err
    kind, cfg = FULL_MODELS["st_pgat_spgnn_3"]
    g = synth_device.make_batch(0, 6, ragged=True).graph
    spe.distance_pos_enc(g, pos_enc_dim=39)
    cw = torch.tensor(runner.CLASS_WEIGHTS_22, device="cuda")

    def fresh(drop):
        torch.manual_seed(0)
        net = sm.GATPositionSPGNNNet(**dict(cfg, feat_drop=drop, attn_drop=drop)).cuda()
        net.init()
        net.train()
        net.set_gcn_only()
        ops.manual_seed(3)
        return net, runner.FlatSGD(net.parameters(), lr=5e-3, momentum=0.9)

    # (a) no randomness: eager 3 warm-up + 4 steps == graphed (3 warm-up inside) + 4 replays
    net_e, opt_e = fresh(0.0)
    eager = [float(runner.train_step(net_e, g, opt_e, cw, 1.0).item()) for _ in range(7)]
    net_g, opt_g = fresh(0.0)
    gs = runner.GraphedTrainStep(net_g, g, opt_g, cw, 1.0, warmup=3)
    replayed = [float(gs().item()) for _ in range(4)]
    gs.reset_salt()
    assert np.allclose(replayed, eager[3:], rtol=2e-4), (replayed, eager)         # replay k is training step 3 + k
    assert rel_err(opt_g.flat_p.cpu(), opt_e.flat_p.cpu()) < 1e-4
    # (b) with dropout: deterministic, and masks are redrawn per replay
    seqs = []
    for _ in range(2):
        net_d, opt_d = fresh(0.3)
        gd = runner.GraphedTrainStep(net_d, g, opt_d, cw, 0.15, warmup=2)
        seqs.append([float(gd().item()) for _ in range(6)])
        gd.reset_salt()
    # same masks on both runs (only the atomics of the bias-gradient sums reorder)
    assert np.allclose(seqs[0], seqs[1], rtol=1e-5) and np.isfinite(seqs[0]).all()
    assert int(gd.counter.item()) == 0 and gd.replays == 6
