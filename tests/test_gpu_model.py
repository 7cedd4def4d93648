"""End-to-end parity of the accelerated meta-archs against the fixtures the UNMODIFIED reference wrote
(tests/golden/model_*.npz via oracle/make_golden.py): dense logits / offsets / video_cls and the final
segment sets for hard and soft NMS, in fp32 mode (1e-4) and bf16 mode (1e-2), plus batch invariance and the
raw-stream entry point. Tolerances are BASELINE.json's."""
import os

import numpy as np
import pytest
import torch

from make_golden import MODEL_CASES, VIDEO_CASES, make_item
from audio_visual_deepfake_detection_b200.libs.core import load_config_for
from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build(case, precision, **over):
    model_name, overrides, use_video, wseed = MODEL_CASES[case]
    cfg = load_config_for(model_name, dict(overrides, **over))
    model = make_meta_arch(cfg["model_name"], **cfg["model"], precision=precision, max_batch=8)
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed))
    return model.to("cuda").eval(), use_video


def max_rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def match_segments(got_s, got_p, ref_s, ref_p, tol_t, tol_p, final_thr=0.2):
    """Both lists are sorted by score. Same count and scores everywhere; the segments that survive the
    downstream `score > 0.2` filter (generate_results.ipynb cell 2; the north-star bar) must agree to tol_t.
    Below that filter greedy NMS is discontinuous (two overlapping candidates whose decayed scores tie to 1e-6
    may be picked in either order), so there up to max(2, 5 %) of the entries may differ."""
    assert len(got_p) == len(ref_p), (len(got_p), len(ref_p))
    np.testing.assert_allclose(got_p, ref_p, atol=tol_p)
    got_s, ref_s = got_s.reshape(-1, 2), ref_s.reshape(-1, 2)
    keep = ref_p > final_thr
    np.testing.assert_allclose(got_s[keep], ref_s[keep], atol=tol_t)
    if (~keep).any():
        ok = np.abs(got_s[~keep] - ref_s[~keep]).max(axis=1) <= tol_t
        assert (~ok).sum() <= max(2, 0.05 * ok.size), (~ok).sum()


@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_fp32_mode_vs_reference(case):
    model, use_video = build(case, "fp32")
    g = np.load(os.path.join(GOLD, f"model_{case}.npz"))
    items = [make_item(dur, seed, mode, use_video) for dur, seed, mode in VIDEO_CASES]
    logits, offsets, vcls = model.dense_outputs(items)
    for vi in range(len(items)):
        assert max_rel(logits[vi].numpy(), g[f"v{vi}_logits"]) < 1e-4, vi
        assert max_rel(offsets[vi].numpy(), g[f"v{vi}_offsets"]) < 1e-4, vi
        np.testing.assert_allclose(vcls[vi].numpy(), g[f"v{vi}_video_cls"][0], atol=1e-4, rtol=1e-4)
    for method in ("hard", "soft"):
        model.test_nms_method = method
        out = model(items)
        for vi, r in enumerate(out):
            assert r["video_id"] == items[vi]["video_id"]
            match_segments(r["segments"].numpy(), r["scores"].numpy(), g[f"v{vi}_{method}_segments"],
                           g[f"v{vi}_{method}_scores"], 1e-3, 1e-4)
            assert r["labels"].dtype == torch.long and r["video_cls"].shape == (1,)


# "mixed" (the default: bf16 raw features, fp16 bounded activations) must meet the north-star 1e-2. (Pure bf16 operands
# sit at 1.2e-2 on these worst-case synthetic weights - as the reference does under autocast(bf16) - and are not offered.)
@pytest.mark.parametrize("precision,bar", [("mixed", 1e-2)])
@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_bf16_mode_vs_reference(case, precision, bar):
    model, use_video = build(case, precision)
    g = np.load(os.path.join(GOLD, f"model_{case}.npz"))
    items = [make_item(dur, seed, mode, use_video) for dur, seed, mode in VIDEO_CASES]
    logits, offsets, vcls = model.dense_outputs(items)
    worst = 0.0
    for vi in range(len(items)):
        e1 = max_rel(logits[vi].numpy(), g[f"v{vi}_logits"]); e2 = max_rel(offsets[vi].numpy(), g[f"v{vi}_offsets"])
        worst = max(worst, e1, e2)
        assert e1 < bar and e2 < bar, (vi, e1, e2)
        np.testing.assert_allclose(vcls[vi].numpy(), g[f"v{vi}_video_cls"][0], atol=2e-2, rtol=2e-2)
    print("worst max-rel error", precision, case, worst)
    # the final-set bar (identical membership after the 0.2 filter, start/end within 1e-3 s) is asserted at the benchmarked
    # batch size over hundreds of videos in tests/test_gpu_parity_sets.py; here: the post-processing of these four videos
    # is exact on the path's own dense outputs, hard and soft
    import model_ref
    import parity_common
    model_name, overrides, _, wseed = MODEL_CASES[case]
    cfg = load_config_for(model_name, dict(overrides))
    om = model_ref.OracleModel(cfg["model"], syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed), model_name)
    for method in ("hard", "soft"):
        model.test_nms_method = method
        out = model(items)
        for vi, r in enumerate(out):
            L = model.engine().padded_len(int(items[vi]["feats"].shape[-1]))
            want_s, want_p = parity_common.oracle_from_dense(om, logits[vi].numpy(), offsets[vi].numpy(), model.engine().level_lens(L),
                                                             items[vi], method)
            assert len(want_p) == r["scores"].numel(), (case, vi, method)
            np.testing.assert_allclose(r["scores"].numpy(), want_p, atol=2e-6)
            np.testing.assert_allclose(r["segments"].numpy().reshape(-1, 2), want_s, atol=1e-4)


def test_bf16_operand_format_is_not_offered():
    with pytest.raises(ValueError):
        build("exp12", "bf16")[0].engine()


def test_batch_invariance_and_streams():
    model, use_video = build("exp12", "mixed")
    durs = [4.03, 9.04, 26.37, 7.42, 5.5]
    raw = [{"video_id": f"vid{i}", "duration": d, "streams": syn.synthetic_streams(d, 100 + i)} for i, d in enumerate(durs)]
    import interp_ref
    items = [interp_ref.dataset_item(r["streams"], r["duration"], r["video_id"]) for r in raw]
    one_by_one = [model([it])[0] for it in items]
    batched = model(items)
    from_streams = model.forward_streams(raw)
    for a, b, c in zip(one_by_one, batched, from_streams):
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["segments"], b["segments"])
        assert torch.equal(a["video_cls"], b["video_cls"])
        assert torch.equal(a["scores"], c["scores"]) and torch.equal(a["segments"], c["segments"])
        assert a["video_id"] == c["video_id"]


def test_stream_from_pinned_arrays_skips_the_staging_copy():
    """model.stream on page-locked source arrays (avdf_host_all_pinned -> avdf_h2d_gather: the copy engine reads the caller's
    memory) returns exactly what the staging path (avdf_host_pack into pinned slots) returns, batch after batch."""
    model, use_video = build("exp12", "mixed")
    rng = np.random.RandomState(5)
    batches = []
    for bi in range(5):                                             # ragged batch sizes, more batches than would fit one slot
        durs = rng.uniform(4.1, 14.0, size=[3, 1, 4, 2, 3][bi])
        batches.append([{"video_id": f"b{bi}v{i}", "duration": float(d), "streams": syn.synthetic_streams(float(d), 1000 + 10 * bi + i)}
                        for i, d in enumerate(durs)])
    pinned = []
    for chunk in batches:
        new_chunk = []
        for c in chunk:
            st = {}
            for k, a in c["streams"].items():
                t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
                t.numpy()[...] = a
                st[k] = t.numpy() if k != "emo" else t              # numpy views and torch tensors both
            new_chunk.append({**c, "streams": st})
        pinned.append(new_chunk)
    runner = model.runner()
    n0 = runner.n_direct
    want = [r for out in model.stream(iter(batches)) for r in out]
    assert runner.n_direct == n0                                    # pageable arrays: staged
    got = [r for out in model.stream(iter(pinned)) for r in out]
    assert runner.n_direct == n0 + len(batches)                     # pinned arrays: direct
    assert [r["video_id"] for r in got] == [r["video_id"] for r in want]
    for a, b in zip(want, got):
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["segments"], b["segments"]) and torch.equal(a["video_cls"], b["video_cls"])


def test_bf16_feature_shards_and_collated_pinned_batches():
    """The opt-in 16-bit feature-shard format (streaming.bf16_shard: bf16 bits in uint16 arrays) through model.stream, both
    staged (pageable arrays) and direct (one pinned block per batch, collated stream-major: collate_pinned, one merged
    host->device copy per stream): the GPU resampling of bf16 inputs is bit-identical to the fp32-input kernel fed the
    same rounded values, and the end-to-end results are those of model() on the bf16-rounded features."""
    from audio_visual_deepfake_detection_b200 import ops
    from audio_visual_deepfake_detection_b200.libs.modeling.streaming import bf16_shard, collate_pinned
    import interp_ref
    model, use_video = build("exp12", "mixed")
    rng = np.random.RandomState(8)
    batches = []
    for bi in range(3):
        durs = rng.uniform(4.1, 14.0, size=[3, 8, 2][bi])
        batches.append([{"video_id": f"s{bi}v{i}", "duration": float(d), "streams": syn.synthetic_streams(float(d), 2000 + 10 * bi + i)}
                        for i, d in enumerate(durs)])
    shard = [[{**c, "streams": {k: bf16_shard(a) for k, a in c["streams"].items()}} for c in chunk] for chunk in batches]
    rounded = [[{**c, "streams": {k: torch.from_numpy(a).to(torch.bfloat16).float().numpy() for k, a in c["streams"].items()}}
                for c in chunk] for chunk in batches]
    # kernel level: bf16 in == fp32 in on the rounded values
    pk16, pk32 = model.stage(model.pack_streams(shard[1])), model.stage(model.pack_streams(rounded[1]))
    assert pk16["streams"][1].dtype == torch.bfloat16 and model.h2d_bytes(pk16) < 0.55 * model.h2d_bytes(pk32)
    B, L, C = len(shard[1]), model.max_seq_len, model.engine().c_in
    xa = torch.zeros((B, L, C), dtype=torch.bfloat16, device="cuda"); xb = torch.zeros_like(xa)
    ops.interp_concat(pk16["streams"], pk16["offs"], L, xa)
    ops.interp_concat(pk32["streams"], pk32["offs"], L, xb)
    assert torch.equal(xa, xb)
    # end to end: staged shards, direct (collated pinned) shards and fp32 streams of the rounded values agree exactly
    runner = model.runner()
    want = [r for out in model.stream(iter(rounded)) for r in out]
    n0 = runner.n_direct
    got_staged = [r for out in model.stream(iter(shard)) for r in out]
    assert runner.n_direct == n0
    got_direct = [r for out in model.stream(iter([collate_pinned(c) for c in shard])) for r in out]
    assert runner.n_direct == n0 + len(shard)
    for a, b, c in zip(want, got_staged, got_direct):
        assert a["video_id"] == b["video_id"] == c["video_id"]
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["segments"], b["segments"])
        assert torch.equal(a["scores"], c["scores"]) and torch.equal(a["segments"], c["segments"])
    # and they stay at the mixed-precision bar against the fp32 features (the reference's input): dense outputs
    items32 = [interp_ref.dataset_item(c["streams"], c["duration"], c["video_id"]) for c in batches[1]]
    items16 = [interp_ref.dataset_item(c["streams"], c["duration"], c["video_id"]) for c in rounded[1]]
    l32, o32, _ = model.dense_outputs(items32)
    l16, o16, _ = model.dense_outputs(items16)
    assert max_rel(l16.numpy(), l32.numpy()) < 5e-3 and max_rel(o16.numpy(), o32.numpy()) < 5e-3


def test_fused_mlp_path_matches_two_launch_path():
    """engine.fused_mlp routes every block's MLP through avdf_mlp_fused (one launch, hidden activations on chip). Per call
    it agrees with the two-GEMM path to ~1e-6 (tests/test_gpu_kernels.py::test_mlp_fused); over the whole network
    those last-bit differences pass through the 16-bit rounding points of later layers, so the dense outputs are
    compared at the mixed-precision bar. The fused path itself stays batch invariant (bit-exact)."""
    model, _ = build("exp12", "mixed")
    durs = [4.03, 9.04, 26.37]
    import interp_ref
    items = [interp_ref.dataset_item(syn.synthetic_streams(d, 300 + i), d, f"v{i}") for i, d in enumerate(durs)]
    eng = model.engine()
    was = eng.fused_mlp
    try:
        eng.fused_mlp = False
        base = model.dense_outputs(items)
        eng.fused_mlp, eng.fused_mlp_min_rows = True, 0
        fused = model.dense_outputs(items)
        fused_out = model(items)
        single = [model([it])[0] for it in items]
    finally:
        eng.fused_mlp = was
    for a, b in zip(base[:2], fused[:2]):
        for x, y in zip(a, b):
            assert max_rel(y.cpu().numpy(), x.cpu().numpy()) < 5e-3
    for b, c in zip(fused_out, single):
        assert torch.equal(b["scores"], c["scores"]) and torch.equal(b["segments"], c["segments"])


def test_reference_style_driver(tmp_path):
    """inference_one_epoch over a list-of-lists loader writes the reference's JSON records."""
    import json
    from audio_visual_deepfake_detection_b200.libs.utils import inference_one_epoch
    model, use_video = build("exp12", "mixed")
    items = [make_item(dur, seed, mode, use_video) for dur, seed, mode in VIDEO_CASES]
    loader = [[it] for it in items]
    inference_one_epoch(loader, model, -1, output_folder=str(tmp_path))
    rec = json.load(open(tmp_path / "data_left.json"))
    assert [r["video_id"] for r in rec] == [it["video_id"] for it in items]
    assert all(set(r) == {"video_id", "video_cls", "scores", "segments"} for r in rec)


def test_api_edge_cases():
    """Empty list, more videos than max_batch (chunked), hard-NMS config, batched_nms front-end incl. multiclass."""
    from audio_visual_deepfake_detection_b200.libs.utils import batched_nms
    import nms_ref
    from make_golden import sweep_inputs
    model, use_video = build("exp12", "mixed")            # max_batch = 8
    assert model([]) == []
    durs = [4.5 + 0.7 * i for i in range(11)]
    import interp_ref
    items = [interp_ref.dataset_item(syn.synthetic_streams(d, 300 + i), d, f"e{i}") for i, d in enumerate(durs)]
    out = model(items)                                     # 11 > max_batch: two chunks
    assert [r["video_id"] for r in out] == [it["video_id"] for it in items]
    solo = model([items[9]])[0]
    assert torch.equal(solo["scores"], out[9]["scores"]) and torch.equal(solo["segments"], out[9]["segments"])
    assert all(r["scores"].numel() <= model.test_max_seg_num for r in out)
    assert all(bool((r["segments"] >= 0).all()) and bool((r["segments"][:, 1] <= d + 1e-4).all()) for r, d in zip(out, durs))
    # batched_nms with the reference's signature, CPU tensors in / out
    segs, sc = sweep_inputs(1512, 7000 + 1512)
    s, p = torch.from_numpy(segs), torch.from_numpy(sc)
    for soft in (False, True):
        got = batched_nms(s, p, torch.zeros(1512, dtype=torch.long), 0.1, 0.2, 100, use_soft_nms=soft, multiclass=False,
                          sigma=0.75, voting_thresh=0.9)
        want = nms_ref.batched_nms(s, p, torch.zeros(1512, dtype=torch.long), 0.1, 0.2, 100, use_soft_nms=soft,
                                   multiclass=False, sigma=0.75, voting_thresh=0.9)
        assert got[0].device.type == "cpu" and got[2].dtype == torch.long
        assert torch.equal(got[1], want[1])
        np.testing.assert_allclose(got[0].numpy(), want[0].numpy().reshape(-1, 2), atol=1e-4)
    labels = torch.from_numpy((np.arange(1512) % 3).astype(np.int64))
    for soft in (False, True):          # class-agnostic NMS over several classes: labels follow the picks (nms.py:159-180)
        got = batched_nms(s, p, labels, 0.1, 0.2, 100, use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.9)
        want = nms_ref.batched_nms(s, p, labels, 0.1, 0.2, 100, use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.9)
        assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]) and len(set(got[2].tolist())) > 1
        np.testing.assert_allclose(got[0].numpy(), want[0].numpy().reshape(-1, 2), atol=1e-4)
    got = batched_nms(s, p, labels, 0.1, 0.2, 100, use_soft_nms=False, multiclass=True, sigma=0.75, voting_thresh=0.9)
    want = nms_ref.batched_nms(s, p, labels, 0.1, 0.2, 100, use_soft_nms=False, multiclass=True, sigma=0.75, voting_thresh=0.9)
    assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2])
    np.testing.assert_allclose(got[0].numpy(), want[0].numpy().reshape(-1, 2), atol=1e-5)


def test_nms_method_none_and_config_keyed_graphs():
    """test_cfg nms_method 'none' (av_fd_no_recon.py:847-858): every decoded candidate, in decode order, converted to
    seconds. And the raw-stream path re-captures its CUDA graph when a test_* attribute changes (the reference reads them
    at call time): hard -> soft -> none -> hard through the same runner give each method's own result."""
    import interp_ref, model_ref, nms_ref
    model, use_video = build("exp12", "mixed")
    durs = [4.03, 9.04, 7.42]
    raw = [{"video_id": f"n{i}", "duration": d, "streams": syn.synthetic_streams(d, 700 + i)} for i, d in enumerate(durs)]
    items = [interp_ref.dataset_item(r["streams"], r["duration"], r["video_id"]) for r in raw]
    res = {}
    for method in ("hard", "soft", "none", "hard"):
        model.test_nms_method = method
        a = model.forward_streams(raw)
        b = model(items)
        for x, y in zip(a, b):
            assert torch.equal(x["scores"], y["scores"]) and torch.equal(x["segments"], y["segments"])
        if method in res:                                    # back to 'hard': the cached graph of that config
            for x, y in zip(a, res[method]):
                assert torch.equal(x["scores"], y["scores"]) and torch.equal(x["segments"], y["segments"])
        res[method] = a
    assert not torch.equal(res["hard"][0]["scores"], res["soft"][0]["scores"][: len(res["hard"][0]["scores"])]) or \
        len(res["hard"][0]["scores"]) != len(res["soft"][0]["scores"])
    # 'none' against the oracle's decode applied to the CUDA path's own dense outputs
    logits, offsets, _ = model.dense_outputs(items)
    model_name, overrides, _, wseed = MODEL_CASES["exp12"]
    cfg = load_config_for(model_name, dict(overrides, **{"test_cfg.nms_method": "none"}))
    om = model_ref.OracleModel(cfg["model"], syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed), model_name)
    lens = model.engine().level_lens(768)
    for vi, it in enumerate(items):
        lg, of, ms, o = [], [], [], 0
        for l, n in enumerate(lens):
            lg.append(logits[vi, o:o + n].reshape(n, 1)); of.append(offsets[vi, o:o + n]); ms.append(torch.ones(n, dtype=torch.bool)); o += n
        segs, scores, labels = om.decode(lg, of, ms)
        segs, scores, labels = om.postprocess(segs, scores, labels, it, nms_ref.batched_nms)
        got = res["none"][vi]
        assert got["scores"].numel() == scores.numel() > 100
        np.testing.assert_allclose(got["scores"].numpy(), scores.numpy(), atol=2e-6)
        np.testing.assert_allclose(got["segments"].numpy(), segs.numpy(), atol=1e-4)


def test_result_record_ring():
    """avdf_postprocess appends one fixed-size record per video ([index, count, video_cls, scores[K], segs[K][2]]) to a
    device ring: the unit the multi-GPU gather moves. Rows are self-describing, order is arbitrary."""
    import interp_ref
    model, use_video = build("exp12", "mixed")
    model.test_nms_method = "soft"
    durs = [4.03, 9.04, 26.37, 7.42, 5.5]
    raw = [{"video_id": f"r{i}", "duration": d, "streams": syn.synthetic_streams(d, 800 + i)} for i, d in enumerate(durs)]
    K = model.test_max_seg_num
    ring = torch.zeros((16, 3 + 3 * K), dtype=torch.float32, device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    staged = model.stage(model.pack_streams(raw))
    staged["records"] = (ring, counter)
    staged["vidx"] = torch.tensor([40, 41, 42, 43, 44], dtype=torch.int32, device="cuda")
    res = model.run_staged(staged)
    want = model.fetch(res)
    torch.cuda.synchronize()
    assert int(counter.item()) == len(durs)
    rows = ring[: len(durs)].cpu()
    assert sorted(rows[:, 0].tolist()) == [40.0, 41.0, 42.0, 43.0, 44.0]
    for row in rows:
        w = want[int(row[0]) - 40]
        n = int(row[1])
        assert n == w["scores"].numel()
        assert torch.equal(row[3:3 + n], w["scores"]) and torch.equal(row[3 + K:3 + K + 2 * n].reshape(n, 2), w["segments"])
        assert float(row[2]) == float(w["video_cls"][0])
        assert float(row[3 + n:3 + K].abs().sum()) == 0.0


def test_model_on_second_device_matches_first():
    """`devices: ['cuda:1']` (what the reference yaml ships is a non-zero ordinal): every launch, tensor map, stream and the
    per-device kernel attributes follow the MODEL's device, not the process' current one - same records as on cuda:0, with
    cuda:0 left current on the calling thread; the BYOL-A extractor likewise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    assert torch.cuda.current_device() == 0
    durs = [4.03, 9.04, 7.42]
    raw = [{"video_id": f"vid{i}", "duration": d, "streams": syn.synthetic_streams(d, 300 + i)} for i, d in enumerate(durs)]
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        model_name, overrides, _, wseed = MODEL_CASES["exp12"]
        cfg = load_config_for(model_name, dict(overrides))
        model = make_meta_arch(cfg["model_name"], **cfg["model"], precision="mixed", max_batch=4)
        model.load_state_dict(syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed))
        model.to(dev).eval()
        outs.append(model.forward_streams(raw))
        assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["segments"], b["segments"]) and torch.equal(a["video_cls"], b["video_cls"])
    from audio_visual_deepfake_detection_b200.libs.features import AudioNTT2020Task6
    wavs = [syn.synthetic_wav(16000 * 2 + 77, 5), syn.synthetic_wav(16000 + 999, 6)]
    feats = []
    for dev in ("cuda:0", "cuda:1"):
        m = AudioNTT2020Task6().load_state_dict(syn.synthetic_byola_state_dict(0)).to(dev).eval()
        feats.append([f.cpu() for f in m.extract(wavs)])
    for a, b in zip(*feats):
        assert torch.equal(a, b)
