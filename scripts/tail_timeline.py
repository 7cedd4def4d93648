"""Debug: globaltimer stamps of the block-tail kernel's first tile on CTA 0 (projection + LN2 prologue, the eight hidden
chunks, final epilogue)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_visual_deepfake_detection_b200 import ops, native as nv
dev = "cuda"
L = nv.lib()
L.avdf_debug_mlp_timeline.argtypes = [ctypes.c_void_p]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 148
C, H, dt = 256, 1024, torch.float16
att = torch.randn(rows, C, device=dev).to(dt); wo = (torch.randn(C, C, device=dev) / 16).to(dt)
w1 = (torch.randn(H, C, device=dev) / 16).to(dt); w2 = (torch.randn(C, H, device=dev) / 32).to(dt)
b1 = torch.randn(H, device=dev); b2 = torch.randn(C, device=dev); gam = torch.ones(C, device=dev); bo = torch.randn(C, device=dev)
ln2 = (torch.ones(C, device=dev), torch.zeros(C, device=dev))
skip = torch.randn(rows, C, device=dev); y = torch.empty(rows, C, device=dev)
mask = torch.ones(rows, dtype=torch.uint8, device=dev); out = torch.empty(rows, C, device=dev)
call = lambda: ops.mlp_fused(None, w1, b1, w2, b2, row_mask=mask, residual=None, gamma=None, out=out, proj=(att, wo, bo, gam, ln2, skip, None))
call(); torch.cuda.synchronize()
dbg = torch.zeros(288, dtype=torch.int64, device=dev)
L.avdf_debug_mlp_timeline(ctypes.c_void_p(dbg.data_ptr()))
call(); torch.cuda.synchronize()
L.avdf_debug_mlp_timeline(None)
d = dbg.cpu().view(3, 96)
t0 = int(d[d > 0].min())
rel = lambda v: None if v == 0 else int(v) - t0
print("producer: att issue", rel(d[0, 0]), "slot issue times:", [rel(v) for v in d[0, 1:41].tolist()])
print("epilogue: skip requested", rel(d[2, 42]), " proj acc ready", rel(d[2, 43]), " pass 1 done", rel(d[2, 44]), " pass 2 done", rel(d[2, 45]))
print("mma: x_full", rel(d[1, 0]), "G1(0) issued", rel(d[1, 1]))
for j in range(8):
    print("  j=%d  G1(j+1) issued %s  hidden ready %s  G2 issued %s | epi: acc1 full %s  math done %s  published %s" %
          (j, rel(d[1, 2 + 4 * j]), rel(d[1, 3 + 4 * j]), rel(d[1, 4 + 4 * j]), rel(d[2, 4 * j]), rel(d[2, 4 * j + 1]), rel(d[2, 4 * j + 3])))
print("final: acc2 full", rel(d[2, 40]), "stores read", rel(d[2, 41]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    call(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(20):
            call()
g.replay(); torch.cuda.synchronize(); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("rows %d: %.2f us per launch (graph of 20)" % (rows, 1000 * e0.elapsed_time(e1) / 20))
