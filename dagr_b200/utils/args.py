"""argparse + yaml flag surface (reference: src/dagr/utils/args.py:7-110): CLI wins, yaml fills the rest."""
import argparse
from pathlib import Path

import yaml


def BASE_FLAGS():
    p = argparse.ArgumentParser("")
    S = argparse.SUPPRESS
    for name, typ in (("dataset_directory", Path), ("output_directory", Path), ("checkpoint", Path), ("img_net", str),
                      ("img_net_checkpoint", Path), ("task", str), ("dataset", str), ("radius", float),
                      ("time_window_us", int), ("max_neighbors", int), ("n_nodes", int), ("batch_size", int),
                      ("activation", str), ("edge_attr_dim", int), ("aggr", str), ("kernel_size", int),
                      ("pooling_aggr", str), ("base_width", float), ("after_pool_width", float),
                      ("net_stem_width", float), ("yolo_stem_width", float), ("num_scales", int),
                      ("pooling_dim_at_output", str), ("weight_decay", float), ("clip", float), ("aug_p_flip", float)):
        p.add_argument("--" + name, type=typ, default=S)
    p.add_argument("--config", type=Path, default="../config/detection.yaml")
    for flag in ("use_image", "no_events", "pretrain_cnn", "keep_temporal_ordering"):
        p.add_argument("--" + flag, action="store_true")
    return p


def parse_config(args, config: Path):
    with Path(config).open() as f:
        cfg = yaml.load(f, Loader=yaml.SafeLoader)
    for k, v in cfg.items():
        if k not in args:
            setattr(args, k, v)
    return args


def FLAGS(argv=None):
    p = BASE_FLAGS()
    S = argparse.SUPPRESS
    for name, typ in (("aug_trans", float), ("aug_zoom", float), ("exp_name", str), ("l_r", float), ("tot_num_epochs", int)):
        p.add_argument("--" + name, type=typ, default=S)
    p.add_argument("--no_eval", action="store_true")
    p.add_argument("--run_test", action="store_true")
    p.add_argument("--num_interframe_steps", type=int, default=10)
    args = p.parse_args(argv)
    if str(args.config) != "":
        args = parse_config(args, args.config)
    return args


def default_args(size: str = "s", dataset: str = "dsec", **over):
    """Namespace equal to config/dagr-{n,s,m,l}-dsec.yaml / dagr-l-ncaltech.yaml (reference config/*.yaml)."""
    w = dict(n=0.25, s=0.5, m=0.75, l=1.0)[size]
    d = dict(task="detection", dataset=dataset, radius=0.01, time_window_us=1000000, max_neighbors=16, n_nodes=50000,
             batch_size=8, activation="relu", edge_attr_dim=2, aggr="sum", kernel_size=5, pooling_aggr="max",
             base_width=0.5, after_pool_width=1, net_stem_width=w, yolo_stem_width=w, num_scales=2,
             pooling_dim_at_output="5x7", img_net="resnet18", use_image=False, no_events=False, pretrain_cnn=False,
             keep_temporal_ordering=False)
    if dataset == "ncaltech101":
        d.update(num_scales=1)
    d.update(over)
    return argparse.Namespace(**d)
