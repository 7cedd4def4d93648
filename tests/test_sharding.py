"""Host-side multi-rank logic on CPU (gloo, world_size 2): rank-strided sharding + the one all-gather of
fixed-size result records reproduce the single-process result list."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_visual_deepfake_detection_b200.libs.utils import sharding

K = 7
N_VIDEOS = 11


def fake_result(i):
    g = torch.Generator().manual_seed(100 + i)
    n = int(torch.randint(0, K + 3, (1,), generator=g))
    return {"video_id": "v%d" % i, "segments": torch.rand((n, 2), generator=g), "scores": torch.rand(n, generator=g).sort(descending=True).values,
            "video_cls": torch.randn(1, generator=g)}


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sharding.shard_indices(N_VIDEOS, rank, world)
    rec = sharding.pack_records(idx, [fake_result(i) for i in idx], K)
    allrec = sharding.gather_records(rec, N_VIDEOS)
    got = sharding.unpack_records(allrec, K)
    q.put((rank, sorted(got), {k: (v["scores"].tolist(), v["segments"].tolist(), v["video_cls"].tolist()) for k, v in got.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_shards_cover_every_video_once():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in sharding.shard_indices(N_VIDEOS, r, world))
        assert seen == list(range(N_VIDEOS))
        sizes = [len(sharding.shard_indices(N_VIDEOS, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = sharding.unpack_records(sharding.pack_records(list(range(N_VIDEOS)), [fake_result(i) for i in range(N_VIDEOS)], K), K)
    for rank, keys, got in outs:
        assert keys == list(range(N_VIDEOS))
        for i in range(N_VIDEOS):
            n = min(len(fake_result(i)["scores"]), K)
            assert got[i][0] == single[i]["scores"].tolist() and len(got[i][0]) == n
            assert got[i][1] == single[i]["segments"].tolist()
            assert got[i][2] == single[i]["video_cls"].tolist()
