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
    t0 = d[0, 0].item()
    if t0 == 0:
        print(f'--- M={B*T} N={N} K={K} residual={residual} ln={ln} act={act} half_out={half_out}')
    names = ["start", "setup done", "t2 first TMA landed", "t2 MMA committed", "t2 epi got acc", "epilogue done", "after final sync", "-", "t2 epi enter", "t2 vectors", "t2 c0 tmem", "t2 c0 staged", "t2 c0 stored", "t2 c1 tmem", "t2 c1 staged", "t2 c1 stored"]
    t8 = d[0, 8].item()
    if t0 != 0: print(f"--- M={B*T} N={N} K={K} residual={residual} ln={ln} act={act} half_out={half_out}: CTA0 tile-2 phases (ns from epi enter):", {n: int(d[0, i].item() - t8) for i, n in enumerate(names) if n.startswith('t2')}, 'total', int(d[0,6].item()-t0))
    ends = d[:, 6]; used = ends > 0
    print("    all CTAs: start spread %d ns, end-start max %d ns" % (int((d[used, 0].max() - d[used, 0].min()).item()), int((ends[used].max() - d[used, 0].min()).item())))
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
