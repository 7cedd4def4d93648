"""One batch of the BYOL-A extractor (32 clips, AV-Deepfake1M durations) inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off [--set full] python scripts/byola_step.py
Not a benchmark: nothing printed here is a number to report."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_visual_deepfake_detection_b200.libs.features import AudioNTT2020Task6, BatchPlan, LogMelSpectrogram    # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn                                         # noqa: E402

dev = "cuda:0"
m = AudioNTT2020Task6().load_state_dict(syn.synthetic_byola_state_dict(0)).to(dev).eval()
durs = syn.sample_durations(32, seed=4321)
wavs = [syn.synthetic_wav(int(16000 * d), 9000 + i) for i, d in enumerate(durs)]
mel = LogMelSpectrogram(dev)
plan = BatchPlan([w.shape[0] for w in wavs], dev)
wav = torch.from_numpy(np.concatenate(wavs)).to(dev)
for _ in range(2):
    m.forward_packed(mel.packed(wav, plan), plan)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.forward_packed(mel.packed(wav, plan), plan)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
