"""GPU parity of the BYOL-A feature extractor against oracle/byola_ref.py and the fixtures the
reference's own AudioNTT2020Task6 + torchaudio's MelSpectrogram produced (tests/golden/byola.npz). Bars: fp32 mode 1e-4
of the largest feature value, 16-bit operands 1e-2 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import byola_ref
from audio_visual_deepfake_detection_b200 import native as nv
from audio_visual_deepfake_detection_b200 import ops
from audio_visual_deepfake_detection_b200.libs.features import AudioNTT2020Task6, BatchPlan, LogMelSpectrogram
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "byola.npz"))
DEV = "cuda"


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _clips():
    return [syn.synthetic_wav(int(n), int(seed)) for n, seed in GOLD["clips"]]


def _model(precision):
    m = AudioNTT2020Task6(n_mels=64, d=2048, precision=precision)
    m.load_state_dict(syn.synthetic_byola_state_dict(int(GOLD["weight_seed"])))
    return m.to(DEV).eval()


def test_logmel_against_torchaudio_fixture():
    """avdf_logmel on a packed batch against the reference's `normalizer((to_melspec(wav) + eps).log())`."""
    lms = LogMelSpectrogram(DEV)(_clips())
    for i, got in enumerate(lms):
        want = GOLD[f"lms{i}"]
        got = got.cpu().numpy()
        assert got.shape == want.shape
        err = np.abs(got - want)
        assert err.max() < 5e-3                       # bins at the log(eps) floor hold the FFT's fp32 rounding noise
        assert err[want > -1.0].max() < 1e-4          # 85 % of the bins
        assert err[want > 0.0].max() < 2e-5


@pytest.mark.parametrize("mode", ["fp32", "f16", "bf16"])
def test_conv_gemm_tap_table_is_conv2d(mode):
    """avdf_conv_gemm with nine row offsets over the padded grid layout == F.conv2d(padding=1) + bias, ReLU, with the
    outputs at padding rows forced to zero (two 'clips' separated by a zero step)."""
    rng = np.random.RandomState(5)
    adt = {"fp32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[mode]
    mel, lens = 16, [37, 20]
    steps = 64
    x = np.zeros((steps, mel + 2, 64), np.float32)
    mask = np.zeros((steps, mel + 2), np.uint8)
    s, clips = 1, []
    for t in lens:
        x[s:s + t, 1:mel + 1] = rng.standard_normal((t, mel, 64))
        mask[s:s + t, 1:mel + 1] = 1
        clips.append((s, t)); s += t + 1
    w = (rng.standard_normal((64, 64, 3, 3)) / np.sqrt(576)).astype(np.float32)          # [out, in, mel tap, time tap]
    b = rng.normal(0, 0.3, 64).astype(np.float32)
    rows = steps * (mel + 2)
    xd = torch.from_numpy(x).reshape(rows, 64).to(DEV, adt)
    wd = torch.from_numpy(w).permute(0, 3, 2, 1).reshape(64, 576).contiguous().to(DEV, adt)
    out = torch.full((rows, 64), float("nan"), device=DEV, dtype=adt)
    taps = [dt * (mel + 2) + dm for dt in (-1, 0, 1) for dm in (-1, 0, 1)]
    ops.conv_gemm(xd, wd, taps=9, batch=1, c_in=64, n_out=64, segs=[(rows, 0, 0)], a_rows=rows, o_rows=rows,
                  bias=torch.from_numpy(b).to(DEV), row_mask=torch.from_numpy(mask).reshape(-1).to(DEV), act=ops.ACT_RELU, tap_rows=taps,
                  **({"out_f32": out} if mode == "fp32" else {"out_h": out}))
    got = out.float().cpu().reshape(steps, mel + 2, 64)
    assert torch.isfinite(got).all()
    xr = xd.float().cpu().reshape(steps, mel + 2, 64)
    wr = wd.float().cpu().reshape(64, 3, 3, 64).permute(0, 3, 2, 1)                      # the rounded operands the kernel saw
    for s0, t in clips:
        clip = xr[s0:s0 + t, 1:mel + 1].permute(2, 1, 0)[None]                            # (1, ch, mel, time)
        want = F.relu(F.conv2d(clip.double(), wr.double(), torch.from_numpy(b).double(), padding=1))[0].permute(2, 1, 0)
        assert rel_err(got[s0:s0 + t, 1:mel + 1], want) < (2e-5 if mode == "fp32" else 4e-3)
    assert float(got[torch.from_numpy(mask) == 0].abs().max()) == 0.0


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("mixed", 1e-2), ("bf16", 1e-2)])
def test_extract_against_reference_fixture(precision, bar):
    """wav -> features through the public API, the four clips as ONE packed batch, against what the reference's own model
    class returned for every clip alone."""
    m = _model(precision)
    feats = m.extract(_clips())
    torch.cuda.synchronize()
    for i, f in enumerate(feats):
        want = GOLD[f"feat{i}"]
        assert tuple(f.shape) == want.shape
        e = rel_err(f, want)
        print(precision, i, "rel err %.2e" % e)
        assert e < bar, (precision, i, e)
        assert abs(float((f > 0).float().mean()) - float((want > 0).mean())) < 0.01      # the ReLU pattern is alive, not flat


def test_batch_packing_invariance_and_module_forward():
    """A clip's features do not depend on what it is batched with (every clip sees exactly its own zero padding): bit-equal
    alone vs packed; `model(lms)` with the reference's (B, 1, 64, T) input agrees with the oracle on that lms."""
    m = _model("mixed")
    clips = _clips()
    packed = m.extract(clips)
    for i, c in enumerate(clips):
        alone = m.extract([c])[0]
        assert torch.equal(alone, packed[i]), i
    lms = torch.from_numpy(np.stack([GOLD["lms0"][:, :96], GOLD["lms3"][:, 100:196]]))[:, None].to(DEV)     # (2, 1, 64, 96)
    out = m(lms)
    assert tuple(out.shape) == (2, 12, 2048)
    sd = syn.synthetic_byola_state_dict(int(GOLD["weight_seed"]))
    for b in range(2):
        assert rel_err(out[b], byola_ref.forward(lms[b, 0].cpu().numpy(), sd)) < 1e-2
    m32 = _model("fp32")
    out32 = m32([lms[0, 0], lms[1]])
    for b in range(2):
        assert rel_err(out32[b], byola_ref.forward(lms[b, 0].cpu().numpy(), sd)) < 1e-4


def test_edge_cases():
    """The shortest legal clip (1120 samples: 8 frames, one output row), odd frame counts at every level; clips the reference
    cannot process either (no room for the reflect padding, fewer than 8 frames for the three poolings) are refused."""
    m = _model("fp32")
    sd = syn.synthetic_byola_state_dict(int(GOLD["weight_seed"]))
    wavs = [syn.synthetic_wav(n, 40 + i) for i, n in enumerate((1120, 160 * 15 + 1, 160 * 8, 160 * 23 + 159))]
    outs = m.extract(wavs)
    for w, o in zip(wavs, outs):
        want = byola_ref.extract(w, sd)
        assert tuple(o.shape) == want.shape
        assert rel_err(o, want) < 1e-4
    assert outs[0].shape[0] == 1
    for bad in (512, 1119):
        with pytest.raises(ValueError):
            m.extract([np.zeros(bad, np.float32)])
    with pytest.raises(nv.AvdfError):
        ops.conv_gemm(torch.zeros((128, 64), device=DEV, dtype=torch.float16), torch.zeros((64, 640), device=DEV, dtype=torch.float16), taps=10,
                      batch=1, c_in=64, n_out=64, segs=[(128, 0, 0)], a_rows=128, o_rows=128, tap_rows=list(range(10)),
                      out_h=torch.zeros((128, 64), device=DEV, dtype=torch.float16))


def test_extractor_feeds_the_localization_model():
    """The extractor's output IS the model's BYOL-A stream (deepfake_video_audio.py:451-479: the `.npy` the dataset loads and
    truncates): wav -> GPU extractor -> raw-stream entry point of the localization model. The same pipeline fed with the CPU
    oracle's features gives the same video-level logit to within the 16-bit operand noise."""
    from audio_visual_deepfake_detection_b200.libs.core import load_config_for
    from audio_visual_deepfake_detection_b200.libs.modeling import EXP12, make_meta_arch
    cfg = load_config_for(EXP12, {"dataset.video_input_dim": 0, "test_cfg.nms_method": "soft"})
    model = make_meta_arch(cfg["model_name"], **cfg["model"], max_batch=4)
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], EXP12, seed=3))
    model.to(DEV).eval()
    ext = _model("mixed")
    durs = [5.0, 7.3]
    wavs = [syn.synthetic_wav(int(16000 * d), 70 + i) for i, d in enumerate(durs)]
    feats = ext.extract(wavs)
    raw = []
    for i, d in enumerate(durs):
        st = syn.synthetic_streams(d, 500 + i, video_dim=0)
        t_b = st["byola"].shape[0]
        assert feats[i].shape[0] >= t_b                      # 12.5 rows per second; the dataset truncates to int(12.497 d - 0.3657)
        st["byola"] = feats[i][:t_b].cpu().numpy()
        raw.append({"video_id": "v%d" % i, "duration": d, "streams": st})
    out = model.forward_streams(raw)
    assert len(out) == 2 and all(o["segments"].shape[1] == 2 for o in out)
    assert all(np.isfinite(np.asarray(o["scores"])).all() for o in out)
    sd = syn.synthetic_byola_state_dict(int(GOLD["weight_seed"]))
    raw_ref = []
    for r, w in zip(raw, wavs):
        st = dict(r["streams"])
        st["byola"] = byola_ref.extract(w, sd)[:st["byola"].shape[0]]
        raw_ref.append(dict(r, streams=st))
    out_ref = model.forward_streams(raw_ref)
    for a, b in zip(out, out_ref):
        va, vb = float(np.asarray(a["video_cls"]).reshape(-1)[0]), float(np.asarray(b["video_cls"]).reshape(-1)[0])
        assert abs(va - vb) < 2e-2 * max(1.0, abs(vb)), (va, vb)
