"""Debug: the NMS sweep of bench.py (BASELINE.json configs[3]) on its own: GPU time of avdf_nms_hard / avdf_nms_soft at
N = 1k ... 100k next to the CPU port of nms_cpu.cpp, bit-equality of the pick lists. AVDF_NMS_CLUSTER=0 runs lists longer
than the shared-memory capacity on one CTA (the round-1 path)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch          # noqa: E402
import bench          # noqa: E402

torch.cuda.set_device(0)
print(json.dumps(bench.nms_sweep(torch.device("cuda:0")), indent=1))
