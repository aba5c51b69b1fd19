#!/usr/bin/env python
"""GNN-stage training entry point — same CLI as /root/reference/train.py:24-41.

    python train.py [lr] [pretrain] --smp <exp_settings file | preset name> [--cpk_path P] [--batch_size B]
                    [--synthetic N] [--epochs E] [--steps S]

``--smp`` takes a reference-style settings file unchanged, or a preset name (``st_pgat_spgnn_3``, ``st_gat_3``, …).
``--synthetic N`` trains on N synthetic airway trees instead of ``DB_PATH`` pickles.
"""
import logging
from argparse import ArgumentParser

from spgnn_b200.settings import Settings, get_callable_by_name


def run_training_job(args):
    settings = Settings(args.smp or "st_pgat_spgnn_3")
    settings.OPTIMIZER = dict(settings.OPTIMIZER, lr=args.lr)
    settings.RELOAD_CHECKPOINT_PATH = args.cpk_path
    if args.batch_size > 0:
        settings.TRAIN_BATCH_SIZE = args.batch_size
    settings.RELOAD_CHECKPOINT = args.pretrain > 0
    if args.synthetic:
        settings.SYNTHETIC_SCANS = args.synthetic
    ct = get_callable_by_name(settings.JOB_RUNNER_CLS)(settings)
    return ct.run(max_epochs=args.epochs, steps=args.steps)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(message)s")
    parser = ArgumentParser()
    parser.add_argument("lr", type=float, nargs="?", default=5e-4, help="set up learning rate.")
    parser.add_argument("pretrain", type=int, nargs="?", default=0, help="if use pretrained model.")
    parser.add_argument("--smp", type=str, nargs="?", default=None, help="settings module path or preset name.")
    parser.add_argument("--cpk_path", type=str, default=None, help="set checkpoint path.")
    parser.add_argument("--batch_size", type=int, default=0, help="scans per batch.")
    parser.add_argument("--synthetic", type=int, default=0, help="use N synthetic airway trees.")
    parser.add_argument("--epochs", type=int, default=None, help="override NUM_EPOCHS.")
    parser.add_argument("--steps", type=int, default=None, help="override GCN_STEPS.")
    run_training_job(parser.parse_args())
