"""Debug: banded attention launch cost (CUDA-graph replay of 20 launches) per pyramid level, stacked q/k/v layout of the
engine. AVDF_ATT_MMA=0 selects the CUDA-core kernel, default the mma.sync kernel (csrc/attention_mma.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_visual_deepfake_detection_b200 import ops

dev = "cuda"


def timed(call, reps=20):
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, 1000 * e0.elapsed_time(e1) / reps)
    return best


for T in (768, 384, 192, 96, 48):
    B, C = 32, 256
    qkv = torch.randn(B, 3 * T, C, device=dev).to(torch.float16)
    mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
    out = torch.empty(B, T, C, device=dev, dtype=torch.float16)
    us = timed(lambda: ops.attention(None, None, None, mask, out, batch=B, t=T, n_head=4, window=7, qkv=qkv))
    print("T=%4d  %.2f us per launch (AVDF_ATT_MMA=%s)" % (T, us, os.environ.get("AVDF_ATT_MMA", "1")), flush=True)
