#!/usr/bin/env python
"""GNN-stage inference entry point — same CLI as /root/reference/test.py:25-37.

    python test.py --smp <exp_settings file | preset name> [--output_path DIR] [--ckp_path P] [--synthetic N]
"""
import logging
from argparse import ArgumentParser

from spgnn_b200.settings import Settings, get_callable_by_name


def run_testing_job(args):
    settings = Settings(args.smp or "st_gat_3")
    settings.RELOAD_CHECKPOINT_PATH = args.ckp_path
    if args.synthetic:
        settings.SYNTHETIC_SCANS = args.synthetic
    ct = get_callable_by_name(settings.TEST_RUNNER_CLS)(output_path=args.output_path, settings_module=settings)
    return ct.run()


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(message)s")
    parser = ArgumentParser()
    parser.add_argument("--smp", type=str, nargs="?", default=None, help="settings module path or preset name.")
    parser.add_argument("--output_path", type=str, nargs="?", default=None, help="set up output path.")
    parser.add_argument("--ckp_path", type=str, nargs="?", default=None, help="checkpoint to load.")
    parser.add_argument("--synthetic", type=int, default=0, help="use N synthetic airway trees.")
    run_testing_job(parser.parse_args())
