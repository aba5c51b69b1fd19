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
