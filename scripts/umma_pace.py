"""Debug: ns per tcgen05.mma (128 x N x 16, fp16, shared-memory operands) with 1 and 148 CTAs issuing."""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch
# the microbenchmark kernel is not part of the product library: built here into scripts/bin/libumma_pace.so
CSRC = os.path.join(os.path.dirname(HERE), "audio_visual_deepfake_detection_b200", "csrc")
SO = os.path.join(HERE, "bin", "libumma_pace.so")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "umma_pace.cu")):
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
                           "-Xcompiler", "-fPIC", "-I", CSRC, os.path.join(HERE, "umma_pace.cu"), os.path.join(CSRC, "runtime.cu"),
                           "-o", SO, "-cudart", "static"])
L = ctypes.CDLL(SO)
L.avdf_debug_umma_pace.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 2000
for ctas in (1, 148):
    for n in (128, 256):
        for two in (0, 6, 8, 14):
            for _ in range(2):
                rc = L.avdf_debug_umma_pace(n, iters, two, ctas, ctypes.c_void_p(out.data_ptr()), None)
                torch.cuda.synchronize()
            ns = out[:ctas].float().mean().item() / (iters * 4)
            tf = 2 * 128 * n * 16 / ns * 1e-3 * ctas
            print("ctas %3d  N=%3d  two_acc=%d: %.1f ns per instruction (%.0f TFLOP/s over %d SMs)" % (ctas, n, two, ns, tf, ctas))
