"""Shared driver of the final-segment-set parity checks (tests/test_gpu_parity_sets.py, scripts/parity_probe.py,
bench.py's cpu_baseline leg): runs N synthetic videos through the CUDA path at the benchmarked batch size and precision,
and through the oracle (oracle/model_ref.py + oracle/nms_ref.c) one by one like the reference (av_fd_no_recon.py:456),
then compares the final sets after the 0.2 score filter. TEST INFRASTRUCTURE."""
import numpy as np
import torch

import interp_ref
import model_ref
import nms_ref
import parity
from make_golden import MODEL_CASES
from audio_visual_deepfake_detection_b200.libs.core import load_config_for
from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn


def make_videos(n_videos, use_video, seed0=5000, include_tiny=True):
    """The 12 tinydataset-shaped clips (BASELINE.json configs[0]) followed by AV-Deepfake1M-length clips."""
    raw = []
    byola = 0 if use_video == "video+emo" else 2048          # the exp5-style case: visual stream + emotion2vec only
    if include_tiny:
        for i in range(min(n_videos, len(syn.TINYDATASET_SHAPES))):
            d, st = syn.tinydataset_streams(i, seed0 + i, video_dim=256 if use_video else 0, byola_dim=byola)
            raw.append({"video_id": "tiny%02d" % i, "duration": d, "streams": st})
    durs = syn.sample_durations(max(0, n_videos - len(raw)), seed=seed0)
    for j, d in enumerate(durs):
        raw.append({"video_id": "syn%04d" % j, "duration": float(d),
                    "streams": syn.synthetic_streams(float(d), seed0 + 100 + j, video_dim=256 if use_video else 0, byola_dim=byola)})
    return raw


def oracle_from_dense(om, logits, offsets, lens, item, method):
    """Oracle decode + NMS + voting + seconds on GIVEN dense outputs (one video)."""
    lg, of, ms, o = [], [], [], 0
    T = int(item["feats"].shape[-1])
    for l, n in enumerate(lens):
        lg.append(torch.as_tensor(logits[o:o + n]).reshape(n, 1))
        of.append(torch.as_tensor(offsets[o:o + n]).reshape(n, 2))
        ms.append((torch.arange(n) * om.fpn_strides[l]) < T)
        o += n
    segs, scores, labels = om.decode(lg, of, ms)
    tc = dict(om.test_cfg); tc["nms_method"] = method
    old = om.test_cfg
    om.test_cfg = tc
    try:
        segs, scores, labels = om.postprocess(segs, scores, labels, item, nms_ref.batched_nms)
    finally:
        om.test_cfg = old
    return segs.numpy().reshape(-1, 2), scores.numpy()


WEIGHTS = {"dense": -3.0, "sparse": -7.0}      # cls prior bias of the synthetic weights (see synthetic_state_dict)


def run_parity(case, n_videos, precision="mixed", batch=32, methods=("hard", "soft"), seed0=5000, via_streams=True,
               weights="dense"):
    model_name, overrides, use_video, wseed = MODEL_CASES[case]
    cfg = load_config_for(model_name, dict(overrides))
    sd = syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed, cls_bias=WEIGHTS[weights])
    model = make_meta_arch(cfg["model_name"], **cfg["model"], precision=precision, max_batch=batch)
    model.load_state_dict(sd)
    model.to("cuda").eval()
    om = model_ref.OracleModel(cfg["model"], sd, model_name)
    raw = make_videos(n_videos, use_video, seed0)
    items = [interp_ref.dataset_item(r["streams"], r["duration"], r["video_id"]) for r in raw]
    eng = model.engine()
    lens, strides = eng.level_lens(eng.max_seq_len), eng.strides
    tc = dict(cfg["model"]["test_cfg"])
    # ---- CUDA path, `batch` videos per call
    gl, go, gv = [], [], []
    for i in range(0, n_videos, batch):
        a, b, c = model.dense_outputs(items[i:i + batch])
        gl.append(a.numpy()); go.append(b.numpy()); gv.append(c.numpy())
    gl, go, gv = np.concatenate(gl), np.concatenate(go), np.concatenate(gv)
    got = {}
    for m in methods:
        model.test_nms_method = m
        if via_streams:          # the raw-stream entry point bench.py times (interp/concat on the GPU, CUDA graph)
            got[m] = [r for i in range(0, n_videos, batch) for r in model.forward_streams(raw[i:i + batch])]
        else:
            got[m] = model(items)
    # ---- oracle, one video per call like the reference
    stats = {m: {"videos": n_videos, "identical": 0, "membership_diff": 0, "boundary_only": 0, "explained": 0, "unexplained": [],
                 "post_exact_fail": [], "post_max_dscore": 0.0, "post_max_dt": 0.0, "max_dt_matched": 0.0, "n_members_ref": 0, "n_dt_over": 0, "max_dt_over_tol_ratio": 0.0}
             for m in methods}
    dense_err = {"logits": 0.0, "offsets": 0.0, "vcls": 0.0, "score_abs": 0.0}
    for vi, item in enumerate(items):
        x, mask = om.preprocess(item["feats"].to(torch.float32))
        lg, of, masks, vcls = om.forward_dense(x, mask)
        rl = torch.cat([a[0].permute(1, 0).flatten() for a in lg]).numpy()
        ro = torch.cat([a[0].permute(1, 0) for a in of], dim=0).numpy()
        e_l = float(np.abs(gl[vi] - rl).max() / np.abs(rl).max()); e_o = float(np.abs(go[vi] - ro).max() / np.abs(ro).max())
        dense_err["logits"] = max(dense_err["logits"], e_l); dense_err["offsets"] = max(dense_err["offsets"], e_o)
        dense_err["vcls"] = max(dense_err["vcls"], abs(float(gv[vi]) - float(vcls[0])) / max(1.0, abs(float(vcls[0]))))
        sc_r, sg_r = parity.dense_to_points(rl, ro, lens, strides)
        sc_g, sg_g = parity.dense_to_points(gl[vi], go[vi], lens, strides)
        pm = np.concatenate([a[0, 0].numpy() for a in masks]).astype(bool)
        tol_s = float(np.abs(sc_g - sc_r)[pm].max())
        live = pm & (sc_r > tc["pre_nms_thresh"])
        tol_x = float(np.abs(sg_g - sg_r)[live].max()) if live.any() else 0.0
        dense_err["score_abs"] = max(dense_err["score_abs"], tol_s)
        sec_per_unit = item["feat_stride"] / item["fps"]
        for m in methods:
            st = stats[m]
            ref_s, ref_p = oracle_from_dense(om, rl, ro, lens, item, m)
            chk_s, chk_p = oracle_from_dense(om, gl[vi], go[vi], lens, item, m)
            g_s, g_p = got[m][vi]["segments"].numpy().reshape(-1, 2), got[m][vi]["scores"].numpy()
            # (ii) the CUDA post-processing is exact on its own dense outputs
            # (scores to 2e-6: the kernel's sigmoid is 1 / (1 + expf(-x)) with CUDA's expf, torch.sigmoid differs by an ulp)
            if len(g_p) != len(chk_p):
                st["post_exact_fail"].append({"video": item["video_id"], "n": [len(g_p), len(chk_p)]})
            elif len(g_p):
                ds, dt = float(np.abs(g_p - chk_p).max()), float(np.abs(g_s - chk_s).max())
                st["post_max_dscore"] = max(st["post_max_dscore"], ds); st["post_max_dt"] = max(st["post_max_dt"], dt)
                if ds > 2e-6 or dt > 1e-4:
                    st["post_exact_fail"].append({"video": item["video_id"], "dscore": ds, "dt": dt})
            # (iii) final sets vs the reference
            c = parity.compare_sets(g_s, g_p, ref_s, ref_p)
            st["n_members_ref"] += c["n_ref"]
            st["max_dt_matched"] = max(st["max_dt_matched"], c["max_dt"])
            st["n_dt_over"] += c["n_dt_over"]
            if c["same_membership"] and c["n_dt_over"] == 0:
                st["identical"] += 1
                continue
            mg = parity.ref_margins(sc_r, sg_r, pm, m, tc, sec_per_unit, tol_x=tol_x)
            tol_dec = 4.0 * tol_s            # decayed scores: own error + the error of the decay factors
            near = (mg["thr"] <= tol_dec) or (mg["order"] <= 2 * tol_dec) or (mg["iou_ratio"] <= 1.0) or (mg["vote_ratio"] <= 1.0)
            # a boundary that moved by no more than the dense boundary tolerance of this video (in seconds) is the
            # offsets' own tolerance, not a set difference
            within_dense = c["same_membership"] and c["max_dt"] <= 2.0 * tol_x * sec_per_unit + 1e-3
            if c["same_membership"]:
                st["boundary_only"] += 1
                st["max_dt_over_tol_ratio"] = max(st["max_dt_over_tol_ratio"], c["max_dt"] / max(tol_x * sec_per_unit, 1e-9))
            else:
                st["membership_diff"] += 1
            if near or within_dense:
                st["explained"] += 1
            else:
                st["unexplained"].append({"video": item["video_id"], "cmp": c, "margins": {k: float(v) for k, v in mg.items()},
                                          "tol_s": tol_s, "tol_x": tol_x})
    return {"case": case, "weights": weights, "precision": precision, "batch": batch, "dense_err": dense_err, "sets": stats}
