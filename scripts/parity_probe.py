"""Print the final-set parity statistics of tests/test_gpu_parity_sets.py for a few videos (GPU box):
    python scripts/parity_probe.py [n_videos] [case ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity_common  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for case in (sys.argv[2:] or ["audio_only", "exp12", "exp13"]):
    for weights in ("dense", "sparse"):
        print(json.dumps(parity_common.run_parity(case, n, weights=weights)))
        sys.stdout.flush()
