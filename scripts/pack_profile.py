"""Debug: where does the host time of StreamRunner._pack go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn
cfg, name, use_video, _ = bench.build_cfg("audio")
model = make_meta_arch(cfg["model_name"], **cfg["model"], max_batch=32)
model.load_state_dict(syn.synthetic_state_dict(cfg["model"], name, seed=0)); model.to("cuda").eval()
raw = bench.make_raw_batches(4, use_video, 11)
r = model.runner()
for _ in model.stream(raw): pass
slot = r.slots[0]
for nthreads in (1, 2, 4, 8, 16, 32):
    r.n_threads = nthreads
    ts = []
    for i in range(12):
        t0 = time.perf_counter(); r._pack(slot, raw[i % 4]); ts.append(time.perf_counter() - t0)
    print("pack threads", nthreads, "ms", round(1000 * np.median(ts), 2))
# raw memcpy bandwidth, single thread
a = np.concatenate([c["streams"]["emo"] for c in raw[0]]); dst = slot.host[2].numpy()[:a.shape[0]]
t0 = time.perf_counter(); np.copyto(dst, a); print("single-thread copy GB/s", a.nbytes / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter(); r._launch(slot); torch.cuda.synchronize(); print("launch+sync ms", 1000 * (time.perf_counter() - t0))
t0 = time.perf_counter(); out = r._collect(slot); print("collect ms", 1000 * (time.perf_counter() - t0))
x = slot.host[2][:a.shape[0]]; d = slot.dev[2][:a.shape[0]]
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("H2D GB/s", x.numel() * 4 / (time.perf_counter() - t0) / 1e9)
