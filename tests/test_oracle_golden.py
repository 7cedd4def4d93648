"""The oracle (oracle/*.py, oracle/nms_ref.c) against the fixtures the
UNMODIFIED reference produced (tests/golden, written by oracle/make_golden.py)
and, when /root/reference is present, against the live reference."""
import json
import os

import numpy as np
import pytest
import torch

import interp_ref
import nms_ref
import ref_harness
from make_golden import MODEL_CASES, VIDEO_CASES, make_item, sweep_inputs
from model_ref import OracleModel
from audio_visual_deepfake_detection_b200.libs.core import load_config_for
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_nms_known_answers():
    kat = json.load(open(os.path.join(GOLD, "nms_kat.json")))
    assert len(kat) >= 7
    for rec in kat:
        segs = np.array(rec["segs"], np.float32).reshape(-1, 2)
        sc = np.array(rec["scores"], np.float32)
        for thr in (0.1, 0.5):
            assert nms_ref.nms(segs, sc, thr).tolist() == rec[f"nms_thr{thr}"], rec["name"]
        for m in (0, 1, 2):
            inds, dets = nms_ref.softnms(segs, sc, 0.1, 0.75, 0.2, m)
            assert inds.tolist() == rec[f"softnms_m{m}"]["inds"], rec["name"]
            assert np.array_equal(dets, np.array(rec[f"softnms_m{m}"]["dets"], np.float32).reshape(-1, 3))
        for soft in (False, True):
            o = nms_ref.batched_nms(torch.from_numpy(segs), torch.from_numpy(sc), torch.zeros(len(sc), dtype=torch.long),
                                    0.1, 0.2, 100, use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.9)
            g = rec[f"batched_{'soft' if soft else 'hard'}"]
            assert o[1].tolist() == pytest.approx(g["scores"], abs=0)
            np.testing.assert_allclose(o[0].numpy().reshape(-1, 2), np.array(g["segs"], np.float32).reshape(-1, 2), atol=1e-4)


def test_survey_edge_cases():
    # answers recorded from the compiled reference
    segs = np.array([[0, 10], [1, 11], [20, 30], [50, 51]], np.float32)
    sc = np.array([.9, .8, .7, .19], np.float32)
    keep = sc > 0.2
    assert nms_ref.nms(segs[keep], sc[keep], 0.1).tolist() == [0, 2]
    inds, dets = nms_ref.softnms(segs, sc, 0.1, 0.75, 0.2, 2)
    assert inds.tolist() == [0, 2, 1]
    assert dets[:, 2].tolist() == pytest.approx([.9, .7, .3276841], abs=1e-6)
    low = np.array([.19, .1, .05], np.float32)
    assert len(nms_ref._hard(segs[:3], low, np.zeros(3, np.int64), 0.1, 0.2, 100)[0]) == 0
    assert len(nms_ref.softnms(segs[:3], low, 0.1, 0.75, 0.2, 2)[0]) == 1


@pytest.mark.parametrize("n", [1000, 1512, 10000, 100000])
def test_nms_sweep_bit_exact(n):
    g = np.load(os.path.join(GOLD, "nms_sweep.npz"))
    segs, sc = sweep_inputs(n, 7000 + n)
    keep = sc > 0.2
    assert np.array_equal(nms_ref.nms(segs[keep], sc[keep], 0.1), g[f"hard_{n}"])
    if n <= 10000:      # the scalar soft-NMS port needs seconds at 100k; the GPU test covers it against the fixture
        inds, dets = nms_ref.softnms(segs, sc, 0.1, 0.75, 0.2, 2)
        assert np.array_equal(inds, g[f"soft_{n}_inds"])
        assert np.array_equal(dets[:, 2], g[f"soft_{n}_scores"])


def test_interp_bit_exact():
    g = np.load(os.path.join(GOLD, "interp.npz"))
    for i, dur in enumerate(g["durations"]):
        st = syn.synthetic_streams(float(dur), 500 + i)
        for k, v in st.items():
            out = interp_ref.linear_resize_tc(v, 768)
            assert np.array_equal(out.reshape(-1)[g[f"{i}_{k}_pos"]], g[f"{i}_{k}_val"]), (i, k)
            assert abs(out.astype(np.float64).sum() - g[f"{i}_{k}_sum"][0]) < 1e-6 * max(1.0, abs(g[f"{i}_{k}_sum"][0]))


@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_model_oracle_vs_golden(case):
    model_name, overrides, use_video, wseed = MODEL_CASES[case]
    cfg = load_config_for(model_name, overrides)
    sd = syn.synthetic_state_dict(cfg["model"], model_name, seed=wseed)
    om = OracleModel(cfg["model"], sd, model_name)
    g = np.load(os.path.join(GOLD, f"model_{case}.npz"))
    for vi, (dur, seed, mode) in enumerate(VIDEO_CASES):
        item = make_item(dur, seed, mode, use_video)
        for method in ("hard", "soft"):
            om.test_cfg = dict(om.test_cfg, nms_method=method)
            r = om([item], nms_ref.batched_nms, return_dense=True)[0]
            gs, gp = g[f"v{vi}_{method}_segments"], g[f"v{vi}_{method}_scores"]
            assert len(r["scores"]) == len(gp), (case, vi, method)
            np.testing.assert_allclose(r["scores"].numpy(), gp, atol=2e-6)
            np.testing.assert_allclose(r["segments"].numpy().reshape(-1, 2), gs.reshape(-1, 2), atol=1e-3)
        np.testing.assert_allclose(r["dense_logits"].numpy(), g[f"v{vi}_logits"], atol=2e-5, rtol=1e-5)
        np.testing.assert_allclose(r["dense_offsets"].numpy(), g[f"v{vi}_offsets"], atol=2e-5, rtol=1e-5)
        np.testing.assert_allclose(r["video_cls"].numpy(), g[f"v{vi}_video_cls"], atol=2e-5)


@pytest.mark.skipif(not ref_harness.available() or not os.path.isfile(os.path.join(ref_harness.REF_SO_DIR, "nms_1d_cpu.so")),
                    reason="reference tree / compiled nms_1d_cpu not present (GPU box)")
def test_oracle_vs_live_reference_nms():
    ref_harness.import_reference()
    import nms_1d_cpu
    rng = np.random.RandomState(3)
    for n in (0, 1, 2, 3, 17, 333, 2000):
        segs, sc = sweep_inputs(n, 10 + n) if n else (np.zeros((0, 2), np.float32), np.zeros(0, np.float32))
        thr = float(rng.choice([0.1, 0.3, 0.5]))
        a = nms_1d_cpu.nms(torch.from_numpy(segs), torch.from_numpy(sc), thr).numpy()
        assert np.array_equal(a, nms_ref.nms(segs, sc, thr))
        for method in (0, 1, 2):
            dets = torch.zeros(max(n, 1), 3)
            a = nms_1d_cpu.softnms(torch.from_numpy(segs), torch.from_numpy(sc), dets, thr, 0.75, 0.2, method).numpy()
            b, d = nms_ref.softnms(segs, sc, thr, 0.75, 0.2, method)
            assert np.array_equal(a, b)
            assert np.array_equal(dets.numpy()[:len(a)], d)


# ---------------------------------------------------------------------------- BYOL-A extractor
def test_byola_oracle_against_reference_fixtures():
    """oracle/byola_ref.py against tests/golden/byola.npz (the reference's AudioNTT2020Task6 class + torchaudio's
    MelSpectrogram, run by oracle/make_golden_byola.py): the mel filters bit-equal, the log-mel spectrogram to its fp32
    noise floor, the features within 5e-5 of the largest value."""
    import byola_ref
    g = np.load(os.path.join(GOLD, "byola.npz"))
    assert np.array_equal(byola_ref.mel_filterbank(), g["mel_fb"])
    sd = syn.synthetic_byola_state_dict(int(g["weight_seed"]))
    for i, (n, seed) in enumerate(g["clips"]):
        wav = syn.synthetic_wav(int(n), int(seed))
        lms = byola_ref.log_mel(wav)
        assert lms.shape == g[f"lms{i}"].shape == (64, byola_ref.n_frames(int(n)))
        # bins without signal sit at log(eps): there the value is the FFT's rounding noise (fp32 in torch, fp64 in numpy)
        assert np.abs(lms - g[f"lms{i}"]).max() < 2e-3
        loud = g[f"lms{i}"] > -1.0            # 85 % of the bins; the error grows as a bin's energy falls towards the noise floor
        assert np.abs(lms - g[f"lms{i}"])[loud].max() < 2e-5
        feat = byola_ref.forward(g[f"lms{i}"], sd)
        assert feat.shape == g[f"feat{i}"].shape == (byola_ref.n_frames(int(n)) // 8, 2048)
        assert np.abs(feat - g[f"feat{i}"]).max() < 2e-5 * np.abs(g[f"feat{i}"]).max()
        full = byola_ref.extract(wav, sd)
        assert np.abs(full - g[f"feat{i}"]).max() < 5e-5 * np.abs(g[f"feat{i}"]).max()


def test_byola_host_side_plan_and_filterbank():
    """BatchPlan's packed layouts (csrc/byola.cu grid layout) and the product's mel filter table, no GPU."""
    import byola_ref
    from audio_visual_deepfake_detection_b200.libs.features import byola
    assert np.array_equal(byola.mel_filterbank(), byola_ref.mel_filterbank())
    plan = byola.BatchPlan([32037, 17123, 9000, 1120], "cpu")
    assert plan.frames == [201, 108, 57, 8] and plan.t[3] == [25, 13, 7, 1]
    assert plan.steps1 % 64 == 0 and plan.steps2 % 64 == 0 and plan.rows3 % 128 == 0
    c1, t1, s1 = plan.d_clip1.numpy(), plan.d_t1.numpy(), plan.d_start1.numpy()
    for c, t in enumerate(plan.t[1]):
        assert np.array_equal(c1[s1[c]:s1[c] + t], np.full(t, c)) and np.array_equal(t1[s1[c]:s1[c] + t], np.arange(t))
        assert c1[s1[c] - 1] == -1 and c1[s1[c] + t] == -1          # a zero step on both sides of every clip
    assert plan.row_off3.tolist() == [0, 25, 38, 45, 46]
    for bad in ([512], [1119]):                   # no reflect padding / fewer than 8 frames: torch refuses both too
        with pytest.raises(ValueError):
            byola.BatchPlan(bad, "cpu")
    assert byola.BatchPlan(None, "cpu", frames=[96, 96]).t[3] == [12, 12]
    m = byola.AudioNTT2020Task6(n_mels=64, d=2048)
    with pytest.raises(KeyError):
        m.load_state_dict({"fc.0.weight": torch.zeros(2048, 512)})
    with pytest.raises(RuntimeError):
        m.load_state_dict(syn.synthetic_byola_state_dict(0)).to("cpu")
    # load_weight's key filter (models.py:27-33): prefixes in front of features. / fc. are dropped, other keys ignored
    sd = {"model.encoder." + k: v for k, v in syn.synthetic_byola_state_dict(0).items()}
    sd["projector.weight"] = torch.zeros(3)
    m2 = byola.AudioNTT2020Task6()
    with pytest.raises(RuntimeError):            # filtered and loaded, then refuses the CPU device
        m2.load_weight(None, "cpu", state_dict={"state_dict": sd})
    assert sorted(m2.state_dict()) == sorted(syn.synthetic_byola_state_dict(0))
