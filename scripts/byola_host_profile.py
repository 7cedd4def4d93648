"""Where the host time of AudioNTT2020Task6.extract() goes (cProfile over 16 batches of 32 clips)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_visual_deepfake_detection_b200.libs.features import AudioNTT2020Task6      # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn          # noqa: E402

m = AudioNTT2020Task6().load_state_dict(syn.synthetic_byola_state_dict(0)).to("cuda:0").eval()
durs = syn.sample_durations(8 * 32, seed=4321)
rng = np.random.RandomState(0)
batches = [[rng.standard_normal(int(16000 * d)).astype(np.float32) for d in durs[32 * b:32 * b + 32]] for b in range(8)]
for bt in batches:
    m.extract(bt)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2):
    for bt in batches:
        m.extract(bt)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
