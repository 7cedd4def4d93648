"""One eager step of the bench workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off [--set full -k regex:... --launch-skip N --launch-count M] python scripts/ncu_step.py
Not a benchmark: nothing printed here is a number to report."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch     # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn      # noqa: E402
import bench                                                                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="audio", choices=list(bench.WORKLOADS))
    ap.add_argument("--precision", default="mixed")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    cfg, name, use_video, _ = bench.build_cfg(args.workload)
    model = make_meta_arch(cfg["model_name"], **cfg["model"], precision=args.precision, max_batch=bench.BATCH)
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], name, seed=0))
    model.to("cuda:0").eval()
    raw = bench.make_raw_batches(1, use_video, seed0=11)[0]
    staged = model.stage(model.pack_streams(raw))
    for _ in range(2):
        model.run_staged(staged)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.run_staged(staged)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
