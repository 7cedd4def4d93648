"""Debug: back-to-back cost (CUDA graph of 20 launches) of the fused MLP kernel against the two-GEMM path."""
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
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / reps


for rows in (32 * 768, 32 * 384, 32 * 96, 32 * 24):
    C, H, dt = 256, 1024, torch.float16
    x = torch.randn(rows, C, device=dev).to(dt); w1 = (torch.randn(H, C, device=dev) / 16).to(dt); w2 = (torch.randn(C, H, device=dev) / 32).to(dt)
    b1 = torch.randn(H, device=dev); b2 = torch.randn(C, device=dev); gam = torch.ones(C, device=dev)
    res = torch.randn(rows, C, device=dev); mask = torch.ones(rows, dtype=torch.uint8, device=dev)
    out = torch.empty(rows, C, device=dev); hid = torch.empty(1, rows, H, device=dev, dtype=dt); out2 = torch.empty(1, rows, C, device=dev)
    fused = lambda: ops.mlp_fused(x, w1, b1, w2, b2, row_mask=mask, residual=res, gamma=gam, out=out)

    def two():
        ops.conv_gemm(x.view(1, rows, C), w1, taps=1, batch=1, c_in=C, n_out=H, segs=[(rows, 0, 0)], a_rows=rows, o_rows=rows, bias=b1, act=ops.ACT_GELU, out_h=hid)
        ops.conv_gemm(hid, w2, taps=1, batch=1, c_in=H, n_out=C, segs=[(rows, 0, 0)], a_rows=rows, o_rows=rows, bias=b2, row_mask=mask.view(1, rows),
                      residual=res.view(1, rows, C), gamma=gam, out_f32=out2)
    print("rows %6d: fused %.2f us, two launches %.2f us, max |diff| %.3e" % (rows, timed(fused), timed(two), float((out - out2.view(rows, C)).abs().max())))
