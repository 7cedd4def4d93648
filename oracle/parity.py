"""Final-segment-set comparison of the CUDA path against the oracle, with the reference-side threshold margins.

TEST INFRASTRUCTURE (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) - never imported by the product.

North-star bar (BASELINE.json): after the downstream `score > 0.2` filter (generate_results.ipynb cell 2) the final
segment sets are identical, start/end within 1e-3 s. The post-processing itself (decode -> NMS -> voting -> seconds,
av_fd_no_recon.py:760-876, libs/utils/nms.py:8-190) is bit-exact on identical inputs, so a set can only differ when the
dense logits / offsets - which agree with the reference to the stated tolerance, not to the bit - move a decision of
that discontinuous pipeline across its threshold. `ref_margins` replays the pipeline on the REFERENCE's dense outputs
and returns how close the reference itself sits to each such decision:

  thr     min |score - min_score| over every score the pipeline compares with min_score: the hard-NMS pre-filter
          (nms.py:15-19), the soft-NMS removal test on every decayed score after every pick (nms_cpu.cpp:139-147), and
          the downstream 0.2 filter on the emitted scores
  order   min score gap of an ordering decision that matters: soft: the picked maximum vs the best remaining candidate it
          overlaps (nms_cpu.cpp:90-106); hard: a kept segment vs a later one it suppresses (nms_cpu.cpp:36-55)
  iou     hard: min | IoU - iou_threshold | over the (kept, later) pairs the greedy loop evaluates
  vote    min | IoU - voting_thresh | over (kept, candidate) pairs whose vote would move a boundary by > 1e-3 s
          (nms.py:67-101)

A mismatching video counts as explained when one of these margins lies within the tolerance measured on that very video
(max |sigma(logit_gpu) - sigma(logit_ref)| for scores, the induced IoU change for the IoU tests).
"""
import numpy as np

F32 = np.float32


def dense_to_points(logits, offsets, lens, strides):
    """Dense head outputs of one video (levels concatenated) -> per-point score [P], seg [P, 2] (feature-grid units),
    exactly the arithmetic of av_fd_no_recon.py:775-812 without the thresholds."""
    logits = np.asarray(logits, F32).reshape(-1)
    offsets = np.asarray(offsets, F32).reshape(-1, 2)
    t, st = [], []
    for n, s in zip(lens, strides):
        t.append(np.arange(n, dtype=F32) * F32(s)); st.append(np.full(n, s, F32))
    t, st = np.concatenate(t), np.concatenate(st)
    score = (F32(1) / (F32(1) + np.exp(-logits.astype(np.float64)))).astype(F32)
    seg = np.stack([t - offsets[:, 0] * st, t + offsets[:, 1] * st], 1).astype(F32)
    return score, seg


def _iou(a, b, eps=0.0):
    """a [2], b [N, 2] -> IoU [N]; eps = 1e-6 in the NMS loops (nms_cpu.cpp:30), 0 in seg_voting (nms.py:84-93)."""
    inter = np.maximum(0.0, np.minimum(a[1], b[:, 1]) - np.maximum(a[0], b[:, 0]))
    la, lb = (a[1] - a[0]) + eps, (b[:, 1] - b[:, 0]) + eps
    return inter / np.maximum(la + lb - inter, 1e-12)


def ref_margins(score, seg, mask, method, tc, sec_per_unit, tol_x=0.0):
    """Replay decode + NMS + voting on one video's per-point reference outputs (float64 arithmetic: margins, not bits).
    Returns dict(thr, order, iou, vote). `tol_x`: measured boundary tolerance (grid units) used to turn IoU margins of the
    pairs into comparable numbers (returned as margin / tolerance ratios under 'iou_ratio', 'vote_ratio')."""
    score = np.asarray(score, np.float64); seg = np.asarray(seg, np.float64)
    ok = (np.asarray(mask) > 0) & (score > tc["pre_nms_thresh"]) & ((seg[:, 1] - seg[:, 0]) > tc["duration_thresh"])
    s, g = score[ok], seg[ok]
    ms, thr_iou, sigma, K = tc["min_score"], tc["iou_threshold"], tc["nms_sigma"], int(tc["max_seg_num"])
    out = {"thr": np.inf, "order": np.inf, "iou": np.inf, "iou_ratio": np.inf, "vote": np.inf, "vote_ratio": np.inf}
    if len(s) == 0:
        return out
    kept_seg, kept_sc = [], []

    def iou_tol(a, b):             # |dIoU| for boundary moves of tol_x on all four ends of the pair
        la, lb = a[1] - a[0], b[:, 1] - b[:, 0]
        return 4.0 * tol_x / np.maximum(np.minimum(la, lb), 1e-6)

    if method == "soft":
        cur, alive = s.copy(), np.ones(len(s), bool)
        n_pick = 0
        while alive.any() and n_pick < K:
            idx = np.flatnonzero(alive)
            i = idx[np.argmax(cur[idx])]
            alive[i] = False
            rest = np.flatnonzero(alive)
            if n_pick > 0 or cur[i] >= ms:      # the first maximum is emitted whatever its score (nms_cpu.cpp:90-106)
                out["thr"] = min(out["thr"], abs(cur[i] - ms))
            kept_seg.append(g[i]); kept_sc.append(cur[i])
            n_pick += 1
            if len(rest) == 0:
                break
            ov = _iou(g[i], g[rest], 1e-6)
            touching = ov > 0
            if touching.any():
                out["order"] = min(out["order"], float(np.min(cur[i] - cur[rest][touching])))
            cur[rest] = cur[rest] * np.exp(-(ov * ov) / sigma)
            out["thr"] = min(out["thr"], float(np.min(np.abs(cur[rest] - ms))))
            alive[rest[cur[rest] < ms]] = False
    else:
        out["thr"] = float(np.min(np.abs(s - ms)))
        keep = s > ms
        s2, g2 = s[keep], g[keep]
        order = np.argsort(-s2, kind="stable")
        s2, g2 = s2[order], g2[order]
        sup = np.zeros(len(s2), bool)
        for i in range(len(s2)):
            if sup[i]:
                continue
            kept_seg.append(g2[i]); kept_sc.append(s2[i])
            if len(kept_sc) >= K:
                break
            later = np.flatnonzero(~sup[i + 1:]) + i + 1
            if len(later) == 0:
                continue
            ov = _iou(g2[i], g2[later], 1e-6)
            d = np.abs(ov - thr_iou)
            out["iou"] = min(out["iou"], float(d.min()))
            if tol_x > 0:
                out["iou_ratio"] = min(out["iou_ratio"], float(np.min(d / np.maximum(iou_tol(g2[i], g2[later]), 1e-12))))
            hit = ov >= thr_iou
            if hit.any():
                out["order"] = min(out["order"], float(np.min(s2[i] - s2[later][hit])))
            sup[later[hit]] = True
    # voting (class-agnostic path, nms.py:174-180): every decoded candidate votes, pre-filter scores
    vt = tc["voting_thresh"]
    if vt > 0 and not tc.get("multiclass_nms", False):
        for k, ks in enumerate(kept_seg):
            if kept_sc[k] <= 0.2:
                continue
            ov = _iou(ks, g, 0.0)
            w = (ov >= vt) * s * ov
            wsum = w.sum()
            if wsum <= 0:
                continue
            # boundary shift (s) if candidate n joined / left the voting set
            wn = s * ov
            center = (w[:, None] * g).sum(0) / wsum
            shift = np.abs(g - center[None]).max(1) * wn / np.maximum(wsum, 1e-12) * sec_per_unit
            rel = shift > 1e-3
            if rel.any():
                d = np.abs(ov[rel] - vt)
                out["vote"] = min(out["vote"], float(d.min()))
                if tol_x > 0:
                    out["vote_ratio"] = min(out["vote_ratio"], float(np.min(d / np.maximum(iou_tol(ks, g[rel]), 1e-12))))
    return out


def compare_sets(got_segs, got_scores, ref_segs, ref_scores, thr=0.2, tol_t=1e-3):
    """Final sets after the `score > thr` filter. Members are matched one to one by boundary distance.
    Returns dict(n_got, n_ref, same_membership, max_dt (s over matched members), n_dt_over (matched members with a
    boundary off by more than tol_t), max_dscore)."""
    gs = np.asarray(got_segs, np.float64).reshape(-1, 2); gp = np.asarray(got_scores, np.float64).reshape(-1)
    rs = np.asarray(ref_segs, np.float64).reshape(-1, 2); rp = np.asarray(ref_scores, np.float64).reshape(-1)
    gs, gp = gs[gp > thr], gp[gp > thr]
    rs, rp = rs[rp > thr], rp[rp > thr]
    res = {"n_got": len(gp), "n_ref": len(rp), "same_membership": len(gp) == len(rp), "max_dt": 0.0, "n_dt_over": 0,
           "max_dscore": 0.0}
    if len(gp) == 0 or len(rp) == 0:
        return res
    free = np.ones(len(gp), bool)
    coarse = 0
    for k in range(len(rp)):
        if not free.any():
            break
        idx = np.flatnonzero(free)
        d = np.abs(gs[idx] - rs[k][None]).max(1)
        j = idx[np.argmin(d)]
        dt = float(np.abs(gs[j] - rs[k]).max())
        # a "member" is the same detection when its boundaries agree to well under its own length; anything farther is
        # a different segment (membership differs)
        if dt > max(0.25 * (rs[k][1] - rs[k][0]), 10 * tol_t):
            coarse += 1
            continue
        free[j] = False
        res["max_dt"] = max(res["max_dt"], dt)
        res["n_dt_over"] += int(dt > tol_t)
        res["max_dscore"] = max(res["max_dscore"], abs(float(gp[j] - rp[k])))
    if coarse:
        res["same_membership"] = False
    return res
