"""Config loader with the reference's semantics (libs/core/config.py:4-164):
a DEFAULTS tree, a recursive "fill what the yaml leaves out" merge, and the
derived fields copied into cfg['model'] (dataset dims, train_cfg, test_cfg).

Only the keys that reach the inference path keep their meaning here; the
optimiser block is carried for drop-in compatibility of the yaml files.
Defaults that matter for parity (config.py:102-115): nms_method 'hard',
iou_threshold 0.1, pre_nms_thresh 0.001.
"""
import copy
import os

import yaml

DEFAULTS = {
    "init_rand_seed": 1234567891,
    "dataset_name": "epic",
    "devices": ["cuda:0"],
    "train_split": ("training",),
    "val_split": ("validation",),
    "model_name": "LocPointTransformer",
    "dataset": {
        "feat_stride": 16, "num_frames": 32, "default_fps": None,
        "audio_feat_folder": None, "audio_file_ext": None,
        "video_input_dim": 2304, "audio_input_dim": 0, "num_classes": 97,
        "downsample_rate": 1, "max_seq_len": 2304, "trunc_thresh": 0.5,
        "crop_ratio": None, "force_upsampling": False,
    },
    "loader": {"batch_size": 8, "num_workers": 4},
    "model": {
        "backbone_type": "convTransformer", "fpn_type": "identity",
        "backbone_arch": (2, 2, 5), "scale_factor": 2,
        "regression_range": [(0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000)],
        "n_head": 4, "n_mha_win_size": -1, "embd_kernel_size": 3, "embd_dim": 512,
        "embd_with_ln": True, "fpn_dim": 512, "fpn_with_ln": True, "fpn_start_level": 0,
        "head_dim": 512, "head_kernel_size": 3, "head_num_layers": 3, "head_with_ln": True,
        "max_buffer_len_factor": 6.0, "use_abs_pe": False, "use_rel_pe": False,
    },
    "train_cfg": {
        "center_sample": "radius", "center_sample_radius": 1.5, "loss_weight": 1.0,
        "cls_prior_prob": 0.01, "init_loss_norm": 2000, "clip_grad_l2norm": -1,
        "head_empty_cls": [], "dropout": 0.0, "droppath": 0.1, "label_smoothing": 0.0,
    },
    "test_cfg": {
        "pre_nms_thresh": 0.001, "pre_nms_topk": 5000, "iou_threshold": 0.1,
        "min_score": 0.01, "max_seg_num": 1000, "nms_method": "hard", "nms_sigma": 0.5,
        "duration_thresh": 0.05, "multiclass_nms": True, "ext_score_file": None,
        "voting_thresh": 0.75,
    },
    "opt": {
        "type": "AdamW", "momentum": 0.9, "weight_decay": 0.0, "learning_rate": 1e-3,
        "epochs": 30, "warmup": True, "warmup_epochs": 5, "schedule_type": "cosine",
        "schedule_steps": [], "schedule_gamma": 0.1,
    },
}

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))))), "configs")


def _fill_missing(defaults, cfg):
    for key, val in defaults.items():
        if key not in cfg:
            cfg[key] = copy.deepcopy(val)
        elif isinstance(val, dict) and isinstance(cfg[key], dict):
            _fill_missing(val, cfg[key])


def _derive(cfg):
    m, d = cfg["model"], cfg["dataset"]
    m["video_input_dim"] = d["video_input_dim"]
    m["audio_input_dim"] = d["audio_input_dim"]
    m["num_classes"] = d["num_classes"]
    m["max_seq_len"] = d["max_seq_len"]
    m["train_cfg"] = cfg["train_cfg"]
    m["test_cfg"] = cfg["test_cfg"]
    return cfg


def load_default_config():
    return copy.deepcopy(DEFAULTS)


def load_config(config_file, defaults=DEFAULTS):
    with open(config_file, "r") as fd:
        cfg = yaml.load(fd, Loader=yaml.FullLoader)
    _fill_missing(defaults, cfg)
    return _derive(cfg)


def load_config_for(model_name, overrides=None):
    """The shipped yaml for a meta-arch name (configs/), with dotted-key
    overrides such as {'dataset.video_input_dim': 0, 'test_cfg.nms_method': 'soft'}."""
    fname = {"AVLocPointTransformerRecoveryNoNormNorecon": "deepfake_exp12_test.yaml",
             "AVLocPointTransformerRecoveryNoNormNoreconTHE": "deepfake_exp13_test.yaml",
             "AVLocPointTransformerRecoveryNoNorm": "deepfake_exp5_test.yaml"}[model_name]
    cfg = load_config(os.path.join(CONFIG_DIR, fname))
    for key, val in (overrides or {}).items():
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = val
    return _derive(cfg)
