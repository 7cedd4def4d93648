"""Time avdf_ln_dwconv_ln (first layout: explicit tile_rows; second layout: tile_rows=0) at the backbone's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_visual_deepfake_detection_b200 import ops
rng = np.random.RandomState(0)
B, C = 32, 256
dev = "cuda"
def params():
    return (torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)).to(dev), torch.from_numpy(rng.normal(0, .1, C).astype(np.float32)).to(dev))
lni = [params() for _ in range(3)]; lno = [params() for _ in range(3)]
dws = [torch.from_numpy(rng.normal(0, 0.6, (C, 3)).astype(np.float32)).to(dev) for _ in range(3)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (T, stride, ns) in ((768, 1, 3), (768, 1, 1), (768, 2, 3), (192, 1, 3)):
    x = torch.randn(B, T, C, device=dev)
    To = T // stride
    mask = torch.ones(B, To, dtype=torch.uint8, device=dev)
    out = torch.empty(B, 3 * To, C, dtype=torch.float16, device=dev)
    skip = torch.empty(B, To, C, device=dev) if stride == 2 else None
    for tile_rows, name in ((8, "v1"), (0, "v2")):
        def run():
            ops.ln_dwconv_ln(x, batch=B, t_src=T, t_virt=T, shift=0, stride=stride, mask_out=mask, ln_in=lni[:ns], dw=dws[:ns],
                             ln_out=lno[:ns], outs=[out] * ns, out_rows=3 * To, out_row_offsets=[0, To, 2 * To][:ns], skip_out=skip,
                             tile_rows=tile_rows)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(60_000_000)          # stall the stream (~30 ms): the launches queue up, the events bracket GPU time only
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        print("T=%d stride=%d streams=%d %s: %.1f us" % (T, stride, ns, name, 1000 * e0.elapsed_time(e1) / reps))
