"""Generate tests/golden/byola.npz by running the UNMODIFIED reference BYOL-A extractor in the build container
(/root/reference must be present). TEST INFRASTRUCTURE.

    python oracle/make_golden_byola.py

What runs: the reference's own `AudioNTT2020Task6` class (audio_feature/content_audio/byol_a/models.py:48-86, loaded from
the file where it lies) and torchaudio's `MelSpectrogram` built with the arguments of
extract_audio_feature_one.py:34-42 / config.yaml, followed by the script's log + PrecomputedNorm line (:66). The
pretrained checkpoint is not shipped, so the weights are the seeded stand-in of
`libs.utils.synthetic.synthetic_byola_state_dict`, loaded through the reference's `load_state_dict`; clips come from
`synthetic_wav` (both regenerated from the seeds stored in the fixture).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn   # noqa: E402

REF = os.environ.get("AVDF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "byola.npz")
# (samples, seed): lengths that exercise odd frame counts at every pooling level and a clip shorter than one second
CLIPS = [(16000 * 2 + 37, 11), (16000 + 160 * 7 + 3, 12), (9000, 13), (16000 * 3 + 1599, 14)]
WEIGHT_SEED = 0


def reference_model():
    spec = importlib.util.spec_from_file_location("ref_byola_models", os.path.join(REF, "audio_feature/content_audio/byol_a/models.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    net = m.AudioNTT2020Task6(n_mels=64, d=2048)          # extract_audio_feature_one.py:44
    sd = syn.synthetic_byola_state_dict(WEIGHT_SEED)
    full = net.state_dict()
    for k, v in sd.items():
        assert full[k].shape == v.shape, k
        full[k] = v
    net.load_state_dict(full)
    return net.eval()


def main():
    import torchaudio
    to_melspec = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160,
                                                      n_mels=64, f_min=60, f_max=7800)
    stats = [-2.2800865, 3.5897882]
    net = reference_model()
    out = {"clips": np.asarray(CLIPS, np.int64), "weight_seed": np.asarray(WEIGHT_SEED),
           "torchaudio_version": np.asarray(torchaudio.__version__), "mel_fb": to_melspec.mel_scale.fb.numpy()}
    for i, (n, seed) in enumerate(CLIPS):
        wav = torch.from_numpy(syn.synthetic_wav(n, seed))[None]
        lms = ((to_melspec(wav) + torch.finfo(torch.float).eps).log() - stats[0]) / stats[1]       # (1, 64, frames)
        with torch.no_grad():
            feats = net(lms.unsqueeze(0))[0]                                                       # (frames // 8, 2048)
        out[f"lms{i}"] = lms[0].numpy()
        out[f"feat{i}"] = feats.numpy()
        print(i, n, tuple(lms.shape), tuple(feats.shape), float(feats.mean()), float((feats > 0).float().mean()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
