import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_gpu_model import build, GOLD, MODEL_CASES, VIDEO_CASES, make_item
for case in ('exp13', 'exp12'):
    model, use_video = build(case, "fp32")
    g = np.load(os.path.join(GOLD, f"model_{case}.npz"))
    items = [make_item(dur, seed, mode, use_video) for dur, seed, mode in VIDEO_CASES]
    for method in ("hard", "soft"):
        model.test_nms_method = method
        out = model(items)
        for vi, r in enumerate(out):
            gs, gp = g[f"v{vi}_{method}_segments"].reshape(-1, 2), g[f"v{vi}_{method}_scores"]
            s, p = r["segments"].numpy(), r["scores"].numpy()
            if len(p) != len(gp):
                print(case, method, vi, "COUNT", len(p), len(gp)); continue
            d = np.abs(s - gs).max(axis=1)
            bad = np.where(d > 1e-3)[0]
            print(case, method, vi, "n", len(p), "max seg diff %.2e score diff %.2e" % (d.max() if len(d) else 0, np.abs(p - gp).max() if len(p) else 0),
                  "bad:", [(int(i), float(p[i]), s[i].tolist(), gs[i].tolist()) for i in bad[:4]])
