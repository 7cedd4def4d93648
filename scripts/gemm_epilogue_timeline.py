"""Debug: globaltimer stamps inside the 16-bit ("wide") epilogue pass of conv_gemm_tc_kernel (library built with
AVDF_GEMM_TIMELINE=1): epilogue warp 4 of every CTA, first tile, its first two 64-column steps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_visual_deepfake_detection_b200 import ops, native as nv

dev = "cuda"
L = nv.lib()
L.avdf_debug_gemm_timeline.argtypes = [ctypes.c_void_p]
B, T, C = 32, 768, 256
a = torch.randn(B, 3 * T, C, device=dev).to(torch.float16); w = (torch.randn(3 * C, C, device=dev) / 16).to(torch.float16)
out = torch.empty(B, 3 * T, C, device=dev, dtype=torch.float16); bias = torch.zeros(3 * C, device=dev)
call = lambda: ops.conv_gemm(a, w, taps=1, batch=B, c_in=C, n_out=C, segs=[(T, i * T, i * T, i * C) for i in range(3)],
                             a_rows=3 * T, o_rows=3 * T, bias=bias, out_h=out)
names = ["kernel start", "setup done", "acc ready", "tmem read", "tile free", "math+sts", "fence", "store issued"]
for mode in (0, 1):
    L.avdf_debug_gemm_ws(mode)
    dbg = torch.zeros(296 * 16, dtype=torch.int64, device=dev)
    call(); torch.cuda.synchronize()
    L.avdf_debug_gemm_timeline(ctypes.c_void_p(dbg.data_ptr()))
    call(); torch.cuda.synchronize()
    L.avdf_debug_gemm_timeline(None)
    d = dbg.cpu().view(296, 16)
    print("mode", "weight-stationary" if mode else "streaming")
    for cta in (0, 1, 77, 147):
        r = d[cta].tolist()
        if r[0] == 0:
            continue
        rel = [int(v - r[0]) if v else None for v in r]
        print("  cta %3d  " % cta + "  ".join("%s %s" % (n, rel[i]) for i, n in enumerate(names)))
        print("           step 2: " + "  ".join("%s %s" % (n, rel[6 + i]) for i, n in enumerate(names) if i >= 2) + "   next tile starts %s" % rel[14])
L.avdf_debug_gemm_ws(-1)
