"""Seeded synthetic weights and AV-Deepfake1M-shaped feature streams.

The reference ships neither a checkpoint nor extracted features
(.MISSING_LARGE_BLOBS), so parity tests and the benchmark run on seeded
stand-ins: numpy RandomState only, so the same seed gives the
same bytes in the build container and on the GPU box.
"""
import math

import numpy as np
import torch

from ..modeling.spec import state_dict_spec

# empirical AV-Deepfake1M test-split duration quantiles (s), from the 343,233
# lines of configs_test/test_folder/*.txt
_DUR_Q = [(0.0, 4.03), (0.05, 5.18), (0.25, 6.02), (0.5, 7.42), (0.75, 10.37), (0.95, 18.75), (0.99, 26.37), (1.0, 33.02)]
BYOLA_FPS = 12.497   # deepfake_video_audio.py:415
EMO_FPS = 50         # deepfake_video_audio.py:416
VIDEO_FPS = 25
# (video_frames, audio_frames @ 16 kHz) of the 12 clips of the reference's tinydataset (tinydataset/metadata/**/*.json:
# BASELINE.json configs[0]); the features themselves are not shipped (.MISSING_LARGE_BLOBS), so the streams of these
# shapes are synthetic too
TINYDATASET_SHAPES = [(240, 154496), (237, 152576), (237, 152576), (240, 154496), (176, 113216), (181, 116736),
                      (181, 116736), (176, 113216), (139, 89216), (134, 86016), (134, 86016), (139, 89216)]


def synthetic_state_dict(model_cfg: dict, model_name: str, seed: int = 0, cls_bias: float = -3.0) -> dict:
    """Every tensor of `state_dict_spec`, non-degenerate: zero-initialised
    tensors of the reference (conv biases, AffineDropPath.scale=1e-4, the
    -4.595 cls prior) are replaced by values that make every path contribute
    and spread the scores across the 0.2 threshold.

    cls_bias = -3.0 (default, the golden fixtures' weights) is a stress case: ~1000 of the 1512 points score above
    0.2, soft-NMS always fills the 100-segment cap. cls_bias = -7.0 ("sparse") gives what a trained detector gives on
    AV-Deepfake1M: a handful of segments per video (the random number stream does not depend on it)."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in state_dict_spec(model_cfg, model_name).items():
        if name.endswith("drop_path_attn.scale") or name.endswith("drop_path_mlp.scale"):
            a = rng.uniform(0.5, 1.5, shape)
        elif name.startswith("reg_head.scale."):
            a = np.asarray(rng.uniform(0.8, 1.6))
        elif ("norm" in name or ".ln" in name or "bn1" in name) and name.endswith(".weight"):
            a = rng.uniform(0.5, 1.5, shape)
        elif ("norm" in name or ".ln" in name or "bn1" in name) and name.endswith(".bias"):
            a = rng.normal(0.0, 0.1, shape)
        elif name == "cls_head.cls_head.conv.bias":
            a = np.full(shape, float(cls_bias))
        elif name.endswith(".bias"):
            a = rng.normal(0.0, 0.1, shape)
        elif name == "cls_head.cls_head.conv.weight":
            a = rng.normal(0.0, 4.0 / math.sqrt(shape[1] * shape[2]), shape)
        elif name == "reg_head.offset_head.conv.weight":
            a = rng.normal(0.0, 1.0 / math.sqrt(shape[1] * shape[2]), shape) + 0.02
        elif name.endswith("_conv.conv.weight") or "fpn_convs" in name:       # depthwise k3
            a = rng.normal(0.0, 0.6, shape)
        else:                                                                   # dense conv / linear
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            a = rng.normal(0.0, 1.0 / math.sqrt(fan_in), shape)
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape).copy())
    return out


def sample_durations(n: int, seed: int = 1234) -> np.ndarray:
    """Inverse-CDF sampling of the empirical duration quantiles above."""
    rng = np.random.RandomState(seed)
    u = rng.uniform(0, 1, n)
    qs = np.array([q for q, _ in _DUR_Q]); vs = np.array([v for _, v in _DUR_Q])
    return np.round(np.interp(u, qs, vs), 2)


def stream_lengths(duration: float):
    """Frames each extractor emits for a clip (deepfake_video_audio.py:461,
    482-483: BYOL-A / emotion2vec streams are truncated to these lengths)."""
    t_v = max(2, int(round(VIDEO_FPS * duration)))
    t_b = max(2, int(BYOLA_FPS * duration - 0.3657))
    t_e = max(2, int(EMO_FPS * duration - 0.817))
    return t_v, t_b, t_e


def synthetic_streams(duration: float, seed: int, video_dim=256, byola_dim=2048, emo_dim=768):
    """Raw per-stream features as the `.npy` files hold them: [T, C] fp32."""
    rng = np.random.RandomState(seed)
    t_v, t_b, t_e = stream_lengths(duration)
    out = {}
    if video_dim:
        out["video"] = rng.standard_normal((t_v, video_dim)).astype(np.float32)
    if byola_dim:
        out["byola"] = np.abs(rng.standard_normal((t_b, byola_dim))).astype(np.float32)
    out["emo"] = rng.standard_normal((t_e, emo_dim)).astype(np.float32)
    return out


def tinydataset_streams(index: int, seed: int, video_dim=256, byola_dim=2048, emo_dim=768):
    """(duration, streams) shaped like clip `index` of the reference's tinydataset: the visual stream has one row per
    video frame, the audio streams the extractors' frame rates over audio_frames / 16 kHz (truncated like
    deepfake_video_audio.py:482-483)."""
    vf, af = TINYDATASET_SHAPES[index]
    duration = af / 16000.0
    rng = np.random.RandomState(seed)
    t_b = max(2, int(BYOLA_FPS * duration - 0.3657))
    t_e = max(2, int(EMO_FPS * duration - 0.817))
    out = {}
    if video_dim:
        out["video"] = rng.standard_normal((vf, video_dim)).astype(np.float32)
    if byola_dim:
        out["byola"] = np.abs(rng.standard_normal((t_b, byola_dim))).astype(np.float32)
    out["emo"] = rng.standard_normal((t_e, emo_dim)).astype(np.float32)
    return duration, out


# ---------------------------------------------------------------------------- BYOL-A extractor
def byola_state_dict_spec(n_mels: int = 64, d: int = 2048) -> dict:
    """Names and shapes of AudioNTT2020Task6's state_dict (audio_feature/content_audio/byol_a/models.py:53-76)."""
    spec = {}
    c_in = 1
    for conv, bn in ((0, 1), (4, 5), (8, 9)):
        spec[f"features.{conv}.weight"] = (64, c_in, 3, 3)
        spec[f"features.{conv}.bias"] = (64,)
        for k in ("weight", "bias", "running_mean", "running_var"):
            spec[f"features.{bn}.{k}"] = (64,)
        c_in = 64
    spec["fc.0.weight"] = (d, 64 * (n_mels // 8)); spec["fc.0.bias"] = (d,)
    spec["fc.3.weight"] = (d, d); spec["fc.3.bias"] = (d,)
    return spec


def synthetic_byola_state_dict(seed: int = 0, n_mels: int = 64, d: int = 2048) -> dict:
    """Seeded stand-in for pretrained_weights/AudioNTT2020-BYOLA-64x96d2048.pth (not shipped): He-scaled dense weights,
    non-trivial BatchNorm statistics, so that every layer keeps O(1) activations with a live ReLU."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in byola_state_dict_spec(n_mels, d).items():
        if name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            a = rng.normal(0.0, 0.2, shape)
        elif len(shape) == 1 and name.endswith(".weight"):          # BatchNorm gamma
            a = rng.uniform(0.5, 1.5, shape)
        elif name.endswith(".bias"):
            a = rng.normal(0.0, 0.1, shape)
        else:
            a = rng.normal(0.0, math.sqrt(2.0 / int(np.prod(shape[1:]))), shape)
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape).copy())
    return out


def synthetic_wav(n_samples: int, seed: int) -> np.ndarray:
    """Speech-like 16 kHz test signal: a few amplitude-modulated partials plus coloured noise, peak below 1 - gives
    the log-mel spectrogram a dynamic range (white noise alone is flat)."""
    rng = np.random.RandomState(seed)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    x = np.zeros(n_samples)
    for _ in range(6):
        f0 = rng.uniform(80, 3500); am = rng.uniform(0.5, 6.0); ph = rng.uniform(0, 2 * np.pi, 2)
        x += rng.uniform(0.02, 0.2) * (0.5 + 0.5 * np.sin(2 * np.pi * am * t + ph[0])) * np.sin(2 * np.pi * f0 * t + ph[1])
    noise = rng.standard_normal(n_samples)
    noise = np.convolve(noise, np.ones(4) / 4.0, mode="same")
    env = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.3, 2.0) * t + rng.uniform(0, 2 * np.pi))
    x += 0.05 * env * noise
    return (x / max(1.0, np.abs(x).max() * 1.05)).astype(np.float32)
