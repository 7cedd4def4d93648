"""CPU restatement of the reference's BYOL-A feature extractor. TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg import this; the product path never does.

What it follows:
  * audio_feature/content_audio/extract_audio_feature_one.py:33-46, 66, 75  (wav -> log-mel -> normalise -> model -> [T, 2048])
  * audio_feature/content_audio/config.yaml:4-9                             (16 kHz, n_fft = win = 1024, hop 160, 64 mels, 60-7800 Hz)
  * audio_feature/content_audio/byol_a/augmentations.py:218-219             (PrecomputedNorm: (x - mean) / std)
  * audio_feature/content_audio/byol_a/models.py:48-86                      (AudioNTT2020Task6: 3 x [conv3x3, BN, ReLU, maxpool 2x2], fc-ReLU-fc-ReLU)
The mel spectrogram itself lives in a third-party dependency that is not vendored in /root/reference: torchaudio
(2.11.0 in this image) `transforms.MelSpectrogram` with its defaults (power 2, centre = True with reflect padding,
periodic Hann window, onesided, HTK mel scale, no filter normalisation). Its published algorithm is restated below
(`mel_filterbank` = torchaudio.functional.melscale_fbanks) and pinned against torchaudio run in the build container
(tests/golden/byola.npz, written by oracle/make_golden_byola.py; tests/test_oracle_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE, N_FFT, HOP, N_MELS, F_MIN, F_MAX = 16000, 1024, 160, 64, 60.0, 7800.0
NORM_STATS = (-2.2800865, 3.5897882)          # extract_audio_feature_one.py:31
EPS = float(np.finfo(np.float32).eps)         # torch.finfo(torch.float).eps, extract_audio_feature_one.py:66


def mel_filterbank(n_freqs=N_FFT // 2 + 1, f_min=F_MIN, f_max=F_MAX, n_mels=N_MELS, sample_rate=SAMPLE_RATE):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> [n_freqs, n_mels] fp32 triangles."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).numpy().astype(np.float32)


def n_frames(n_samples):
    return 1 + n_samples // HOP                # centre = True


def log_mel(wav):
    """wav [n] fp32 at 16 kHz -> normalised log-mel spectrogram [64, frames] fp32."""
    wav = np.asarray(wav, np.float32)
    assert wav.ndim == 1 and wav.shape[0] > N_FFT // 2, "reflect padding needs more than n_fft / 2 samples"
    x = np.pad(wav, (N_FFT // 2, N_FFT // 2), mode="reflect")
    nf = n_frames(wav.shape[0])
    idx = np.arange(nf)[:, None] * HOP + np.arange(N_FFT)[None]
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)).astype(np.float32)     # periodic Hann
    spec = np.fft.rfft((x[idx] * win[None]).astype(np.float32), axis=1)
    power = (spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2).astype(np.float32)   # [frames, 513]
    mel = power @ mel_filterbank()                                                               # [frames, 64]
    lms = (np.log(mel + np.float32(EPS)) - np.float32(NORM_STATS[0])) / np.float32(NORM_STATS[1])
    return np.ascontiguousarray(lms.T.astype(np.float32))


def forward(lms, sd):
    """AudioNTT2020Task6.forward (models.py:78-84) on one clip: lms [64, frames] -> [frames // 8, d] fp32.
    sd: the reference's state_dict keys (features.{0,1,4,5,8,9}.*, fc.{0,3}.*), numpy or torch values."""
    t = lambda k: torch.as_tensor(np.asarray(sd[k]), dtype=torch.float32)
    x = torch.as_tensor(np.asarray(lms), dtype=torch.float32)[None, None]            # (1, 1, mel, time)
    for conv, bn in ((0, 1), (4, 5), (8, 9)):
        x = F.conv2d(x, t(f"features.{conv}.weight"), t(f"features.{conv}.bias"), stride=1, padding=1)
        x = F.batch_norm(x, t(f"features.{bn}.running_mean"), t(f"features.{bn}.running_var"), t(f"features.{bn}.weight"),
                         t(f"features.{bn}.bias"), training=False, eps=1e-5)
        x = F.max_pool2d(F.relu(x), 2, stride=2)
    x = x.permute(0, 3, 2, 1)                                                          # (1, time, mel, ch)
    b, tt, d, c = x.shape
    x = x.reshape(b, tt, c * d)
    x = F.relu(F.linear(x, t("fc.0.weight"), t("fc.0.bias")))                         # Dropout(p=0.3) is the identity in eval()
    x = F.relu(F.linear(x, t("fc.3.weight"), t("fc.3.bias")))
    return x[0].numpy()


def extract(wav, sd):
    return forward(log_mel(wav), sd)
