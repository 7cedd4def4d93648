"""Debug: weight-stationary vs streaming configuration of conv_gemm_tc_kernel on the K = 256 launches of the path
(q/k/v projection, attention projection with residual, FPN lateral), per pyramid level; CUDA-graph replay of 20 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_visual_deepfake_detection_b200 import ops, native as nv

dev = "cuda"
L = nv.lib()


def timed(call, reps=20):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
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


def shape(kind, B, T, dt=torch.float16):
    C = 256
    if kind == "qkv":
        a = torch.randn(B, 3 * T, C, device=dev).to(dt); w = (torch.randn(3 * C, C, device=dev) / 16).to(dt)
        out = torch.empty(B, 3 * T, C, device=dev, dtype=dt); bias = torch.zeros(3 * C, device=dev)
        return lambda: ops.conv_gemm(a, w, taps=1, batch=B, c_in=C, n_out=C, segs=[(T, i * T, i * T, i * C) for i in range(3)],
                                     a_rows=3 * T, o_rows=3 * T, bias=bias, out_h=out)
    a = torch.randn(B, T, C, device=dev).to(dt); w = (torch.randn(C, C, device=dev) / 16).to(dt)
    out = torch.empty(B, T, C, device=dev); bias = torch.zeros(C, device=dev)
    mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
    if kind == "proj":
        res = torch.randn(B, T, C, device=dev); gam = torch.ones(C, device=dev)
        return lambda: ops.conv_gemm(a, w, taps=1, batch=B, c_in=C, n_out=C, segs=[(T, 0, 0)], a_rows=T, o_rows=T, bias=bias,
                                     row_mask=mask, residual=res, gamma=gam, out_f32=out)
    return lambda: ops.conv_gemm(a, w, taps=1, batch=B, c_in=C, n_out=C, segs=[(T, 0, 0)], a_rows=T, o_rows=T, row_mask=mask, out_f32=out)


for kind in ("qkv", "proj", "lateral"):
    for T in (768, 384, 192, 96):
        call = shape(kind, 32, T)
        r = {}
        for mode in (0, 1):
            L.avdf_debug_gemm_ws(mode)
            r[mode] = timed(call)
        L.avdf_debug_gemm_ws(0)
        print("%-8s T=%4d  streaming %6.2f us   weight-stationary %6.2f us" % (kind, T, r[0], r[1]), flush=True)
