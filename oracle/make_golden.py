"""Generate tests/golden/* by running the UNMODIFIED reference in the build
container (/root/reference must be present). TEST INFRASTRUCTURE.

    python oracle/build_ref.py && python oracle/make_golden.py

Fixtures (all small; weights/inputs are regenerated from seeds by
audio_visual_deepfake_detection_b200.libs.utils.synthetic):
  nms_kat.json       known answers of the compiled nms_1d_cpu {nms, softnms}
                     and of libs.utils.batched_nms on hand-made edge cases
  nms_sweep.npz      hard/soft index lists for seeded N in {1000,1512,10000,100000}
  interp.npz         F.interpolate(linear, align_corners=False) samples
  model_<cfg>.npz    reference model(video_list) outputs: dense logits/offsets,
                     video_cls, final segments (hard and soft NMS configs)
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh                                  # noqa: E402
import interp_ref                                         # noqa: E402
from audio_visual_deepfake_detection_b200.libs.modeling.spec import EXP5, EXP12, EXP13   # noqa: E402
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # name: (model_name, cfg overrides, use video stream, weight seed)
    "exp12": (EXP12, {}, True, 0),
    "exp13": (EXP13, {}, True, 0),
    "audio_only": (EXP12, {"dataset.video_input_dim": 0}, False, 3),
    # exp5-style arch with a live reconstruction branch: visual stream + emotion2vec, no BYOL-A
    "exp5": (EXP5, {}, "video+emo", 5),
}
# the reference yaml each case is built from (None: configs_test/deepfake_exp12_test.yaml)
REF_CONFIG = {"exp13": "configs_train/deepfake_exp13.yaml", "exp5": "configs_train/deepfake_exp5.yaml"}
# (duration, seed, mode): 'interp' = dataset path (T=768, all-true mask);
# 'ragged' = feats shorter than max_seq_len fed as-is (masked tail);
VIDEO_CASES = [(4.03, 100, "interp"), (9.04, 101, "interp"), (26.37, 102, "interp"), (7.42, 103, "ragged")]


def sweep_inputs(n, seed):
    rng = np.random.RandomState(seed)
    c = rng.uniform(0, 768, n).astype(np.float32)
    l = rng.uniform(0.01, 40, n).astype(np.float32)
    segs = np.stack([c - l / 2, c + l / 2], 1).astype(np.float32)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    return segs, scores


def make_item(duration, seed, mode, use_video, max_seq_len=768):
    st = syn.synthetic_streams(duration, seed, video_dim=256 if use_video else 0, byola_dim=0 if use_video == "video+emo" else 2048)
    if mode == "interp":
        return interp_ref.dataset_item(st, duration, f"vid{seed}", max_seq_len)
    # ragged: resample every stream to a short common length and do NOT upsample to max_seq_len
    t_short = 500
    item = interp_ref.dataset_item(st, duration, f"vid{seed}", t_short)
    return item


def gen_nms():
    rh.import_reference()
    import nms_1d_cpu
    from libs.utils import batched_nms
    kat = []
    cases = [
        ("survey_mixed", [[0, 10], [1, 11], [20, 30], [50, 51]], [.9, .8, .7, .19]),
        ("all_below_min", [[0, 10], [1, 11], [20, 30]], [.19, .1, .05]),
        ("single", [[3, 4]], [.5]),
        ("ties", [[0, 10], [0, 10], [5, 15], [20, 30], [20, 30]], [.5, .5, .5, .9, .9]),
        ("nested", [[0, 100], [40, 41], [10, 90], [99, 101], [-5, 5]], [.6, .95, .7, .3, .8]),
        ("zero_len", [[5, 5], [5, 5.0005], [4, 6]], [.9, .8, .7]),
        ("empty", [], []),
    ]
    for name, segs, scores in cases:
        s = torch.tensor(segs, dtype=torch.float32).reshape(-1, 2)
        p = torch.tensor(scores, dtype=torch.float32)
        rec = {"name": name, "segs": segs, "scores": scores}
        rec["nms_thr0.1"] = nms_1d_cpu.nms(s, p, 0.1).tolist()
        rec["nms_thr0.5"] = nms_1d_cpu.nms(s, p, 0.5).tolist()
        for method in (0, 1, 2):
            dets = torch.zeros(max(len(scores), 1), 3)
            inds = nms_1d_cpu.softnms(s, p, dets, 0.1, 0.75, 0.2, method)
            rec[f"softnms_m{method}"] = {"inds": inds.tolist(), "dets": dets[:len(inds)].tolist()}
        for soft in (False, True):
            o = batched_nms(s, p, torch.zeros(len(scores), dtype=torch.long), 0.1, 0.2, 100,
                            use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.9)
            rec[f"batched_{'soft' if soft else 'hard'}"] = {"segs": o[0].tolist(), "scores": o[1].tolist()}
        kat.append(rec)
    with open(os.path.join(OUT, "nms_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    sweep = {}
    for n in (1000, 1512, 10000, 100000):
        segs, scores = sweep_inputs(n, 7000 + n)
        s, p = torch.from_numpy(segs), torch.from_numpy(scores)
        keep = p > 0.2
        sweep[f"hard_{n}"] = nms_1d_cpu.nms(s[keep].contiguous(), p[keep].contiguous(), 0.1).numpy().astype(np.int32)
        dets = torch.zeros(n, 3)
        inds = nms_1d_cpu.softnms(s, p, dets, 0.1, 0.75, 0.2, 2)
        sweep[f"soft_{n}_inds"] = inds.numpy().astype(np.int32)
        sweep[f"soft_{n}_scores"] = dets[:len(inds), 2].numpy().copy()
        for soft in (False, True):
            o = batched_nms(s, p, torch.zeros(n, dtype=torch.long), 0.1, 0.2, 100,
                            use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.9)
            sweep[f"batched_{'soft' if soft else 'hard'}_{n}_segs"] = o[0].numpy()
            sweep[f"batched_{'soft' if soft else 'hard'}_{n}_scores"] = o[1].numpy()
        print("nms sweep", n, len(sweep[f"hard_{n}"]), len(inds))
    np.savez_compressed(os.path.join(OUT, "nms_sweep.npz"), **sweep)


def gen_interp():
    out = {}
    rng = np.random.RandomState(99)
    for i, dur in enumerate((4.03, 7.42, 18.75, 33.02)):
        st = syn.synthetic_streams(dur, 500 + i)
        for k, v in st.items():
            ref = F.interpolate(torch.from_numpy(v.T.copy()).unsqueeze(0), size=768, mode="linear",
                                align_corners=False)[0].T.numpy()        # [768, C]
            pos = rng.randint(0, ref.size, 2048)
            out[f"{i}_{k}_pos"] = pos.astype(np.int32)
            out[f"{i}_{k}_val"] = ref.reshape(-1)[pos].copy()
            out[f"{i}_{k}_sum"] = np.array([ref.astype(np.float64).sum()])
    out["durations"] = np.array([4.03, 7.42, 18.75, 33.02])
    np.savez_compressed(os.path.join(OUT, "interp.npz"), **out)


def gen_models():
    for case, (model_name, overrides, use_video, wseed) in MODEL_CASES.items():
        # exp13 has no test yaml in the reference; its own (training) yaml carries the test_cfg
        cfg_path = os.path.join(rh.REF_ROOT, REF_CONFIG[case]) if case in REF_CONFIG else None
        cfg, model = rh.build_reference_model(config_path=cfg_path, model_name=model_name, overrides=overrides)
        sd = syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed)
        model.load_state_dict(sd, strict=True)
        dense = {}

        def hook_cls(m, i, o):
            dense["logits"] = torch.cat([x[0].flatten() for x in o]).clone()

        def hook_reg(m, i, o):
            dense["offsets"] = torch.cat([x[0].permute(1, 0) for x in o], dim=0).clone()
        h1 = model.cls_head.register_forward_hook(hook_cls)
        h2 = model.reg_head.register_forward_hook(hook_reg)
        out = {}
        for vi, (dur, seed, mode) in enumerate(VIDEO_CASES):
            item = make_item(dur, seed, mode, use_video)
            for method in ("hard", "soft"):
                model.test_nms_method = method
                with torch.no_grad():
                    r = model([item])[0]
                out[f"v{vi}_{method}_segments"] = r["segments"].numpy()
                out[f"v{vi}_{method}_scores"] = r["scores"].numpy()
            out[f"v{vi}_video_cls"] = r["video_cls"].numpy()
            out[f"v{vi}_logits"] = dense["logits"].numpy()
            out[f"v{vi}_offsets"] = dense["offsets"].numpy()
            print(case, vi, dur, mode, item["feats"].shape, len(out[f"v{vi}_hard_scores"]), len(out[f"v{vi}_soft_scores"]),
                  float(r["video_cls"]))
        h1.remove(); h2.remove()
        np.savez_compressed(os.path.join(OUT, f"model_{case}.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1:]                 # e.g. `python oracle/make_golden.py exp5`: only these model cases
    if only:
        MODEL_CASES = {k: v for k, v in MODEL_CASES.items() if k in only}
    else:
        gen_nms()
        gen_interp()
    gen_models()
    print("golden fixtures written to", OUT)
