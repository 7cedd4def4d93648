"""Host -> device ceiling of the box, measured with the end-to-end path's own span sizes (SURVEY 8e / VERDICT r1 #6):
every rank copies batches of 32 synthetic videos' raw streams (3 spans per video, ~2.4 MB per video, fp32) from PINNED host
memory, nothing else running on the GPU:
    one   - one cudaMemcpyAsync per batch (the spans packed contiguously)
    spans - avdf_h2d_gather: one cudaMemcpyAsync per span (what model.stream() issues for page-locked inputs)
Run alone or under torchrun (N ranks, one GPU each): rank 0 prints one JSON line with per-rank and aggregate GB/s and the
videos/s those rates would feed. Not a benchmark of the product: a measurement of the box."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_visual_deepfake_detection_b200 import native                      # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, n_batches, seconds = 32, 8, float(os.environ.get("H2D_SECONDS", "1.5"))
    durs = syn.sample_durations(B * n_batches, seed=77 + rank)
    chans = (2048, 768)                                     # audio workload: BYOL-A + emotion2vec
    lens = [syn.stream_lengths(float(d))[1:] for d in durs]
    span_bytes = [[t * c * 4 for t, c in zip(l, chans)] for l in lens]
    per_batch = [sum(sum(sb) for sb in span_bytes[i * B:(i + 1) * B]) for i in range(n_batches)]
    total = sum(per_batch)
    host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = 1                                     # touch: first-touch places the pages on this thread's NUMA node
    dbuf = torch.empty(max(per_batch), dtype=torch.uint8, device=dev)
    L = native.lib()
    st = torch.cuda.Stream()
    res = {}
    for mode in ("one", "spans"):
        # span tables of every batch
        tables = []
        hoff = 0
        for i in range(n_batches):
            src, dst, nb, doff = [], [], [], 0
            for sb in span_bytes[i * B:(i + 1) * B]:
                for x in sb:
                    src.append(host.data_ptr() + hoff + doff); dst.append(dbuf.data_ptr() + doff); nb.append(x); doff += x
            if mode == "one":
                src, dst, nb = [host.data_ptr() + hoff], [dbuf.data_ptr()], [doff]
            n = len(src)
            tables.append(((ctypes.c_void_p * n)(*src), (ctypes.c_void_p * n)(*dst), (ctypes.c_size_t * n)(*nb), n))
            hoff += doff
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            sp = ctypes.c_void_p(st.cuda_stream)
            for t in tables:                                 # warm-up pass
                native.check(L.avdf_h2d_gather(t[0], t[1], t[2], t[3], sp), "avdf_h2d_gather")
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            t0 = time.perf_counter(); moved = 0; k = 0
            while time.perf_counter() - t0 < seconds:
                t = tables[k % n_batches]
                native.check(L.avdf_h2d_gather(t[0], t[1], t[2], t[3], sp), "avdf_h2d_gather")
                moved += per_batch[k % n_batches]; k += 1
                if k % 16 == 0:
                    st.synchronize()                         # keep the queue bounded
            e1.record(st)
            st.synchronize()
        gbs = moved / (e0.elapsed_time(e1) * 1e-3) / 1e9
        res[mode] = {"gbs": gbs, "batches_per_s": k / (e0.elapsed_time(e1) * 1e-3)}
    mine = torch.tensor([res["one"]["gbs"], res["spans"]["gbs"], res["spans"]["batches_per_s"] * B], device=dev)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        a = torch.stack(allr).cpu().numpy()
        numa = {}
        try:
            for nd in sorted(os.listdir("/sys/devices/system/node")):
                if nd.startswith("node"):
                    numa[nd] = open("/sys/devices/system/node/%s/cpulist" % nd).read().strip()
        except OSError:
            pass
        print(json.dumps({"n_gpus": world, "bytes_per_video": total / (B * n_batches), "per_rank_gbs_one_copy_per_batch": [round(float(x), 2) for x in a[:, 0]],
                          "per_rank_gbs_one_copy_per_span": [round(float(x), 2) for x in a[:, 1]], "aggregate_gbs_one": round(float(a[:, 0].sum()), 1),
                          "aggregate_gbs_spans": round(float(a[:, 1].sum()), 1), "videos_per_s_fed_by_spans": round(float(a[:, 2].sum()), 0),
                          "host_cpus": os.cpu_count(), "numa_nodes": numa, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."]}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
