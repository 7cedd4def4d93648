"""Debug: phase timeline of conv_gemm_tc_kernel (globaltimer stamps) and back-to-back launch cost under a CUDA graph."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_visual_deepfake_detection_b200 import ops, native as nv

dev = "cuda"
L = nv.lib()
L.avdf_debug_gemm_timeline.argtypes = [ctypes.c_void_p]

def run(B, T, K, N, residual, ln=False, reps=20, dt=torch.float16, act=0, half_out=False):
    a = torch.randn(B, T, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    out = None if half_out else torch.empty(B, T, N, device=dev); outh = torch.empty(B, T, N, device=dev, dtype=dt) if half_out else None; res = torch.randn(B, T, N, device=dev) if residual else None
    bias = torch.zeros(N, device=dev); mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
    lnp = (torch.ones(N, device=dev), torch.zeros(N, device=dev)) if ln else None
    call = lambda: ops.conv_gemm(a, w, taps=1, batch=B, c_in=K, n_out=N, segs=[(T, 0, 0)], a_rows=T, o_rows=T, bias=bias,
                                 row_mask=None if half_out else mask, residual=res, ln=lnp, act=act, out_f32=out, out_h=outh)
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.avdf_debug_gemm_timeline(ctypes.c_void_p(dbg.data_ptr()))
    call(); torch.cuda.synchronize(); dbg.zero_(); call(); torch.cuda.synchronize()
    d = dbg.cpu().view(148, 16)
    L.avdf_debug_gemm_timeline(None)
    print(f"--- M={B*T} N={N} K={K} residual={residual} ln={ln} act={act} half_out={half_out}")
    if d[0, 0].item() != 0:
        print("    stamps of CTA 0 (ns from the first):", [int(v - d[0, 0].item()) for v in d[0].tolist()])
    # GPU-only cost: graph of `reps` launches
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("    graph replay: %.2f us per launch" % (1000 * e0.elapsed_time(e1) / reps))

run(32, 768, 256, 1024, False, act=2, half_out=True)
run(32, 768, 256, 1024, False, act=0, half_out=True)
run(32, 768, 256, 1024, False, act=0, half_out=False)
run(32, 2304, 256, 256, False, act=0, half_out=True)
run(32, 768, 1024, 256, True)
run(32, 768, 256, 256, True)
