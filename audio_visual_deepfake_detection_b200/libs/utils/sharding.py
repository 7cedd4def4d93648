"""Multi-GPU plumbing of the inference path: videos are independent, so ranks take a strided shard of the video
list (the reference splits its test list into 7 files run as 7 processes, libs/datasets/deepfake_video_audio.py:420-431,
inference.py:121) and the only exchange is one all-gather of fixed-size result records.

record (fp32, width 3 + 3K): [global video index, n_segments, video_cls, scores[K], segments[K, 2] flattened]
"""
import torch
import torch.distributed as dist


def shard_indices(n_videos, rank, world):
    """Rank-strided shard: rank r takes videos r, r+world, ... (balanced to within one video)."""
    return list(range(rank, n_videos, world))


def record_width(K):
    return 3 + 3 * K


def pack_records(indices, results, K, device="cpu"):
    """results: list of model output dicts (segments [N,2], scores [N], video_cls [1]) -> [len, 3+3K] fp32."""
    rec = torch.zeros((len(results), record_width(K)), dtype=torch.float32, device=device)
    for i, (gi, r) in enumerate(zip(indices, results)):
        n = min(int(r["scores"].shape[0]), K)
        rec[i, 0], rec[i, 1], rec[i, 2] = float(gi), float(n), float(r["video_cls"].reshape(-1)[0])
        rec[i, 3:3 + n] = r["scores"][:n]
        rec[i, 3 + K:3 + K + 2 * n] = r["segments"][:n].reshape(-1)
    return rec


def unpack_records(rec, K):
    out = {}
    for row in rec.cpu():
        gi, n = int(row[0]), int(row[1])
        if gi < 0:
            continue
        out[gi] = {"video_cls": row[2:3].clone(), "scores": row[3:3 + n].clone(),
                   "segments": row[3 + K:3 + K + 2 * n].reshape(n, 2).clone()}
    return out


def gather_records(rec, n_videos, world=None):
    """All-gather the per-rank record blocks (padded to the largest shard with index -1 rows)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rec
    world = world or dist.get_world_size()
    per = (n_videos + world - 1) // world
    pad = torch.zeros((per, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    pad[:, 0] = -1
    pad[:rec.shape[0]] = rec
    out = torch.empty((world * per, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, pad)
    return out
