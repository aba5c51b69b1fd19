"""Settings loader and dotted-name resolution for the exp_settings-driven entry points.

Mirrors /root/reference/utils.py:34-61 (``Settings``: a Python file of UPPER_CASE constants) and :522-525
(``get_callable_by_name``).  Reference class paths are mapped onto this package, including the paths that name
classes the reference never shipped: ``apps.airways.labeling_base.job_runner.*`` and the
``*LSPE`` runners / ``models.GATPositionLSPENet`` of st_pgat_spgnnnl_3.py.
"""
from __future__ import annotations

import importlib
import importlib.util

from ._lib import SpgnnError

# GNN hyper-parameters of the shipped exp_settings/*.py files, usable as ``--smp <name>``
_COMMON = dict(out_ch=22, fv_dim=1024, num_hiddens=[256, 128, 64], node_embed_dim=1024)
_GAT = dict(num_heads=2, num_out_heads=2, feat_drop=0.1, attn_drop=0.1, negative_slope=0.2)
PRESETS = {
    "st_gat_3": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATNet", num_gat_layers=3), SAMPLING_RATE=0.3,
                     JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_gat_3_nr": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATNet", num_gat_layers=3, res=False),
                        SAMPLING_RATE=0.3, JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_gat_6": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATNet", num_gat_layers=6,
                                num_hiddens=[256, 128, 64, 64, 64, 64]), SAMPLING_RATE=0.15, NUM_EPOCHS=251,
                     JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_gat_6_nr": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATNet", num_gat_layers=6, res=False,
                                   num_hiddens=[256, 128, 64, 64, 64, 64]), SAMPLING_RATE=0.15, NUM_EPOCHS=251,
                        JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_gcn_3": dict(MODEL=dict(_COMMON, method="models.GCNNet", num_gcn_layers=3), SAMPLING_RATE=0.05,
                     JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_gin_3": dict(MODEL=dict(_COMMON, method="models.GINNet", num_gin_layers=3), SAMPLING_RATE=0.05,
                     JOB_RUNNER_CLS="job_runner.GCNTrain", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_sage_3": dict(MODEL=dict(_COMMON, method="models.SAGENet", num_layers=3, feat_drop=0.1, node_ks=[2, 2, 2, 2],
                                 node_sample_rate=0.3), SAMPLING_RATE=0.05, NODE_BATCH_SIZE=64, TRAIN_BATCH_SIZE=2,
                      JOB_RUNNER_CLS="job_runner.GCNTrainSAGE", TEST_RUNNER_CLS="job_runner.GCNTest"),
    "st_pgat_spgnn_3": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATPositionSPGNNNet", num_gat_layers=3,
                                       pos_hiddens=[256, 128, 64], num_pos_heads=1, pos_enc_dim=39),
                            SAMPLING_RATE=0.15, POS_ENC_DIM=39, JOB_RUNNER_CLS="job_runner.GCNTrainSPGNN",
                            TEST_RUNNER_CLS="job_runner.GCNTestSPGNN"),
    "st_pgat_spgnnnl_3": dict(MODEL=dict(_COMMON, **_GAT, method="models.GATPositionSPGNNNet", num_gat_layers=3,
                                         pos_hiddens=[256, 128, 64], num_pos_heads=1, pos_enc_dim=39, mode="PENL"),
                              SAMPLING_RATE=0.15, POS_ENC_DIM=39, JOB_RUNNER_CLS="job_runner.GCNTrainSPGNN",
                              TEST_RUNNER_CLS="job_runner.GCNTestSPGNN"),
}
DEFAULTS = dict(
    EXP_NAME="spgnn_b200", NR_CLASS=22, TRAIN_BATCH_SIZE=64, VAL_BATCH_SIZE=1, TEST_BATCH_SIZE=1, TRAIN_SAMPLE_SIZE=128,
    GCN_STEPS=300, NUM_EPOCHS=151, SAVE_EPOCHS=50, LOG_STEPS=5, SAMPLING_RATE=0.15, POS_ENC_DIM=39,
    CLASS_WEIGHTS={0: 0.1, 1: 0.2, **{k: 0.8 for k in range(2, 23)}},
    OPTIMIZER={"method": "torch.optim.SGD", "momentum": 0.9, "lr": 1e-4},
    SCHEDULER={"method": "torch.optim.lr_scheduler.ExponentialLR", "gamma": 0.9},
    INITIALIZER={"method": "initializer.HeNorm", "mode": "fan_in"},
    RELOAD_CHECKPOINT=False, RELOAD_CHECKPOINT_PATH=None, RELOAD_DICT_LIST=["model_dict", "metric"],
    DB_PATH=None, MODEL_ROOT_PATH="./models/", USE_DIST_LOSS=False,
)


class Settings:
    """Attributes = the UPPER_CASE names of a settings module (path to a reference-style ``exp_settings/*.py`` file)
    or of a built-in preset name; missing names fall back to the reference's common defaults."""

    def __init__(self, settings_module_path, settings_name="settings"):
        self.settings_module_path = settings_module_path
        self._explicit_settings = set()
        for k, v in DEFAULTS.items():
            setattr(self, k, v.copy() if isinstance(v, (dict, list)) else v)
        if settings_module_path in PRESETS:
            values = {k: (v.copy() if isinstance(v, dict) else v) for k, v in PRESETS[settings_module_path].items()}
        else:
            spec = importlib.util.spec_from_file_location(settings_name, settings_module_path)
            if spec is None:
                raise SpgnnError(f"cannot load settings from {settings_module_path!r} (not a file or preset name; "
                                 f"presets: {sorted(PRESETS)})")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            values = {k: getattr(mod, k) for k in dir(mod) if k.isupper()}
        for k, v in values.items():
            setattr(self, k, v)
            self._explicit_settings.add(k)

    def is_overridden(self, setting):
        return setting in self._explicit_settings

    def __str__(self):
        return "\n".join(f"{k}: {v}" for k, v in sorted(self.__dict__.items()) if k.isupper())


_ALIASES = {
    # reference module paths → this package
    "models.": "spgnn_b200.models.",
    "job_runner.": "spgnn_b200.job_runner.",
    "initializer.": "spgnn_b200.job_runner.",
    "apps.airways.labeling_base.job_runner.": "spgnn_b200.job_runner.",
}
_RENAMES = {
    # names the reference's settings point at but never defined
    "GCNTrainLSPE": "GCNTrainSPGNN", "GCNTestLSPE": "GCNTestSPGNN", "GATPositionLSPENet": "GATPositionSPGNNNet",
}


def get_callable_by_name(module_name):
    mod, _, name = module_name.rpartition(".")
    name = _RENAMES.get(name, name)
    for prefix, target in _ALIASES.items():
        if (mod + ".").startswith(prefix) and not mod.startswith("spgnn_b200"):
            mod = (target + (mod + ".")[len(prefix):]).rstrip(".")
            break
    return getattr(importlib.import_module(mod), name)
