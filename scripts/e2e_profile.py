"""Debug: where the host time of model.stream() goes per batch (pack / launch / collect), audio workload, pinned collated
inputs. Not a benchmark."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch     # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn      # noqa: E402
import bench                                                                      # noqa: E402

torch.cuda.set_device(0)
cfg, name, use_video, _ = bench.build_cfg("audio")
model = make_meta_arch(cfg["model_name"], **cfg["model"], precision="mixed", max_batch=bench.BATCH)
model.load_state_dict(syn.synthetic_state_dict(cfg["model"], name, seed=0))
model.to("cuda:0").eval()
raw = bench.make_raw_batches(8, use_video, seed0=11)
pool = bench.pin_batches(raw)
runner = model.runner()
acc = {"_pack": 0.0, "_launch": 0.0, "_collect": 0.0, "n": 0}
for nm in ("_pack", "_launch", "_collect"):
    fn = getattr(runner, nm)

    def wrap(*a, _fn=fn, _nm=nm, **k):
        t0 = time.perf_counter()
        r = _fn(*a, **k)
        acc[_nm] += time.perf_counter() - t0
        return r
    setattr(runner, nm, wrap)
for _ in model.stream(pool[i % 8] for i in range(40)):
    pass
for k in acc:
    acc[k] = 0
t0 = time.perf_counter()
n = 0
for out in model.stream(pool[i % 8] for i in range(320)):
    n += len(out)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("videos/s %.0f; per batch: wall %.3f ms, pack %.3f, launch %.3f, collect (incl. waiting for the GPU) %.3f" %
      (n / dt, 1e3 * dt / 320, 1e3 * acc["_pack"] / 320, 1e3 * acc["_launch"] / 320, 1e3 * acc["_collect"] / 320))
