"""Host-side rows of SURVEY.md 8(f): `.npy` ingestion (dataset items) and the final result files (notebook logic)."""
import json
import os

import numpy as np
import pytest
import torch

from audio_visual_deepfake_detection_b200.libs.datasets import make_data_loader, make_inference_dataset
from audio_visual_deepfake_detection_b200.libs.utils import merge_results, filter_segments, video_probability
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn


def write_corpus(root, durs, long_audio=True):
    folders = {k: os.path.join(root, k) for k in ("video", "byola", "emo", "lists")}
    for f in folders.values():
        os.makedirs(f, exist_ok=True)
    lines = []
    for i, d in enumerate(durs):
        st = syn.synthetic_streams(d, 40 + i)
        vid = f"id{i:03d}/clip.mp4"
        for k in ("video", "byola", "emo"):
            a = st[k]
            if long_audio and k != "video":           # extractor output is longer than the truncation length
                a = np.concatenate([a, np.ones((5, a.shape[1]), np.float32)], 0)
            path = os.path.join(folders[k], vid.replace(".mp4", ".npy"))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.save(path, a)
        lines.append(f"{vid},{d}")
    with open(os.path.join(folders["lists"], "deepfake_test_sub3.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return folders


def dataset_kwargs(folders):
    return dict(crop_ratio=None, default_fps=None, downsample_rate=0, video_feat_folder=folders["video"],
                audio_feat_folder=None, audio_byola_feat_folder=folders["byola"], audio_emo_feat_folder=folders["emo"],
                audio_file_ext=".npy", num_classes=1, input_dim=0, video_input_dim=256, audio_input_dim=2816, feat_stride=1,
                num_frames=1, test_folder=folders["lists"], trunc_thresh=0.5, max_seq_len=768, force_upsampling=True)


def test_inference_dataset_items(tmp_path):
    durs = [4.03, 7.42, 12.5]
    folders = write_corpus(str(tmp_path), durs)
    ds = make_inference_dataset("deepfake_video_audioEmoBYOLA_inference", False, ["test"], 3, **dataset_kwargs(folders))
    assert len(ds) == 3
    for i, d in enumerate(durs):
        it = ds[i]
        t_v, t_b, t_e = syn.stream_lengths(d)
        assert it["video_id"] == f"id{i:03d}/clip.mp4" and it["duration"] == d
        assert it["streams"]["video"].shape == (t_v, 256)
        assert it["streams"]["byola"].shape == (int(12.497 * d - 0.3657), 2048)      # truncated (:482)
        assert it["streams"]["emo"].shape == (int(50 * d - 0.817), 768)              # truncated (:483)
        assert it["fps"] == pytest.approx(t_v / d)
        assert it["feat_stride"] == pytest.approx(t_v / 768.0) and it["feat_num_frames"] == it["feat_stride"]
    audio = make_inference_dataset("deepfake_audioEmoBYOLA_inference", False, ["test"], 3, **dataset_kwargs(folders))
    assert "video" not in audio[0]["streams"]
    loader = make_data_loader(ds, False, None, 2, 0)
    batches = list(loader)
    assert [len(b) for b in batches] == [2, 1] and isinstance(batches[0], list)
    with pytest.raises(KeyError):
        make_inference_dataset("epic", False, ["test"], 3)


def test_result_files(tmp_path):
    recs_a = [{"video_id": "b.mp4", "video_cls": [3.0], "scores": [0.9, 0.25, 0.1], "segments": [[1, 2], [3, 4], [5, 6]]},
              {"video_id": "a.mp4", "video_cls": [-1.0], "scores": [0.15], "segments": [[0.5, 0.7]]}]
    recs_b = [{"video_id": "b.mp4", "video_cls": [9.0], "scores": [], "segments": []},          # duplicate: first wins
              {"video_id": "c.mp4", "video_cls": [0.0], "scores": [], "segments": []}]
    for name, recs in (("1", recs_a), ("2", recs_b)):
        os.makedirs(tmp_path / name)
        json.dump(recs, open(tmp_path / name / "data_left.json", "w"))
    probs, segs = merge_results([str(tmp_path / "1"), str(tmp_path / "2")], str(tmp_path / "out"))
    assert [p[0] for p in probs] == ["a.mp4", "b.mp4", "c.mp4"]
    assert float(probs[1][1]) == 1.0                                  # sigmoid(3) = 0.953 > 0.9 -> 1.0
    assert float(probs[0][1]) == pytest.approx(1 / (1 + np.e))
    assert segs["b.mp4"] == [[0.9, 1, 2], [0.25, 3, 4]]               # score > 0.2 only
    assert segs["a.mp4"] == [[0, 0, 0]] and segs["c.mp4"] == [[0, 0, 0]]
    assert open(tmp_path / "out" / "prediction.txt").read().splitlines()[1] == "b.mp4;1.0"
    assert json.load(open(tmp_path / "out" / "prediction.json"))["b.mp4"][0] == [0.9, 1, 2]
    assert filter_segments([0.2], [[1, 2]]) == [[0, 0, 0]] and video_probability([0.0]) == 0.5


@pytest.mark.gpu
def test_dataset_to_model_matches_reference_items(tmp_path):
    """dataset item (raw streams, GPU resampling) == the reference's item (CPU F.interpolate restated by the oracle)."""
    import interp_ref
    from audio_visual_deepfake_detection_b200.libs.core import load_config_for
    from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch, EXP12
    durs = [4.03, 9.04, 6.2]
    folders = write_corpus(str(tmp_path), durs, long_audio=False)
    ds = make_inference_dataset("deepfake_video_audioEmoBYOLA_inference", False, ["test"], 3, **dataset_kwargs(folders))
    cfg = load_config_for(EXP12)
    model = make_meta_arch(cfg["model_name"], **cfg["model"], max_batch=4)
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], EXP12, seed=0))
    model.to("cuda").eval()
    items = [ds[i] for i in range(3)]
    got = model(items)
    full = ds.materialize(items[1])
    ref_item = interp_ref.dataset_item(items[1]["streams"], durs[1], items[1]["video_id"])
    assert torch.equal(full["feats"], ref_item["feats"])              # bit-exact interpolation
    want = model([ref_item])[0]
    assert torch.equal(got[1]["scores"], want["scores"]) and torch.equal(got[1]["segments"], want["segments"])


@pytest.mark.gpu
def test_pinned_dataset_items_stream_without_staging(tmp_path):
    """pin_memory=True: the .npy payloads are read straight into page-locked buffers (byte-identical to np.load), and
    model.stream() sends them to the device without the host gather."""
    from audio_visual_deepfake_detection_b200.libs.core import load_config_for
    from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch, EXP12
    durs = [4.03, 9.04, 6.2, 12.5]
    folders = write_corpus(str(tmp_path), durs, long_audio=True)
    plain = make_inference_dataset("deepfake_video_audioEmoBYOLA_inference", False, ["test"], 3, **dataset_kwargs(folders))
    pinned = make_inference_dataset("deepfake_video_audioEmoBYOLA_inference", False, ["test"], 3, pin_memory=True, **dataset_kwargs(folders))
    a, b = [plain[i] for i in range(4)], [pinned[i] for i in range(4)]
    for x, y in zip(a, b):
        for k in x["streams"]:
            assert np.array_equal(x["streams"][k], y["streams"][k]) and y["streams"][k].dtype == np.float32
    cfg = load_config_for(EXP12)
    model = make_meta_arch(cfg["model_name"], **cfg["model"], max_batch=4)
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], EXP12, seed=0))
    model.to("cuda").eval()
    runner = model.runner()
    want = [r for out in model.stream(iter([a])) for r in out]
    n0 = runner.n_direct
    got = [r for out in model.stream(iter([b])) for r in out]
    assert runner.n_direct == n0 + 1
    for u, v in zip(want, got):
        assert torch.equal(u["scores"], v["scores"]) and torch.equal(u["segments"], v["segments"])


@pytest.mark.gpu
def test_inference_cli_end_to_end(tmp_path):
    """`python inference.py cfg sub_index ckpt --merge` on a synthetic corpus: the reference's CLI contract."""
    import subprocess
    import sys
    import yaml
    from audio_visual_deepfake_detection_b200.libs.core import load_config_for
    from audio_visual_deepfake_detection_b200.libs.modeling import EXP12
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    durs = [4.03, 5.5, 7.42, 9.04, 6.2]
    folders = write_corpus(str(tmp_path / "data"), durs)
    cfg = yaml.load(open(os.path.join(root, "configs", "deepfake_exp12_test.yaml")), Loader=yaml.FullLoader)
    cfg["dataset"].update(video_feat_folder=folders["video"], audio_byola_feat_folder=folders["byola"],
                          audio_emo_feat_folder=folders["emo"], test_folder=folders["lists"])
    cfg["loader"] = {"batch_size": 1, "num_workers": 0}
    cfg_path = str(tmp_path / "cfg.yaml")
    yaml.dump(cfg, open(cfg_path, "w"))
    full = load_config_for(EXP12)
    sd = syn.synthetic_state_dict(full["model"], EXP12, seed=0)
    ckpt = str(tmp_path / "run" / "epoch_010.pth.tar")
    os.makedirs(os.path.dirname(ckpt))
    torch.save({"epoch": 10, "state_dict_ema": {"module." + k: v for k, v in sd.items()}}, ckpt)
    r = subprocess.run([sys.executable, os.path.join(root, "inference.py"), cfg_path, "3", ckpt, "-b", "4", "--merge"],
                       capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out_dir = tmp_path / "run" / "3"
    recs = json.load(open(out_dir / "data_left.json"))
    assert [x["video_id"] for x in recs] == [f"id{i:03d}/clip.mp4" for i in range(5)]
    pred = json.load(open(out_dir / "prediction.json"))
    assert set(pred) == {x["video_id"] for x in recs}
    for x in recs:
        kept = [[s, seg[0], seg[1]] for s, seg in zip(x["scores"], x["segments"]) if s > 0.2] or [[0, 0, 0]]
        assert pred[x["video_id"]] == kept
    assert len(open(out_dir / "prediction.txt").read().splitlines()) == 5


def _cli_setup(tmp_path, durs):
    import yaml
    from audio_visual_deepfake_detection_b200.libs.core import load_config_for
    from audio_visual_deepfake_detection_b200.libs.modeling import EXP12
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    folders = write_corpus(str(tmp_path / "data"), durs)
    cfg = yaml.load(open(os.path.join(root, "configs", "deepfake_exp12_test.yaml")), Loader=yaml.FullLoader)
    cfg["dataset"].update(video_feat_folder=folders["video"], audio_byola_feat_folder=folders["byola"],
                          audio_emo_feat_folder=folders["emo"], test_folder=folders["lists"])
    cfg["loader"] = {"batch_size": 1, "num_workers": 0}
    cfg_path = str(tmp_path / "cfg.yaml")
    yaml.dump(cfg, open(cfg_path, "w"))
    full = load_config_for(EXP12)
    sd = syn.synthetic_state_dict(full["model"], EXP12, seed=0)
    return root, folders, cfg_path, full, sd


@pytest.mark.gpu
def test_inference_sharded_single_rank_equals_inference_one_epoch(tmp_path):
    """inference_sharded (records written by the postprocess kernel, gathered, unpacked) == inference_one_epoch, record
    for record, on one rank; partial last batch included."""
    from audio_visual_deepfake_detection_b200.libs.modeling import make_meta_arch
    from audio_visual_deepfake_detection_b200.libs.utils import inference_one_epoch, inference_sharded
    durs = [4.03, 5.5, 7.42, 9.04, 6.2, 11.0, 4.7]
    root, folders, cfg_path, full, sd = _cli_setup(tmp_path, durs)
    ds = make_inference_dataset("deepfake_video_audioEmoBYOLA_inference", False, ["test"], 3, **dataset_kwargs(folders))
    model = make_meta_arch(full["model_name"], **full["model"], max_batch=4)
    model.load_state_dict(sd)
    model.to("cuda").eval()
    a = inference_one_epoch(make_data_loader(ds, False, None, 4, 0), model, -1, output_folder=str(tmp_path / "a"))
    b = inference_sharded(ds, model, str(tmp_path / "b"), batch_size=4)
    assert json.load(open(tmp_path / "a" / "data_left.json")) == json.load(open(tmp_path / "b" / "data_left.json"))
    assert [r["video_id"] for r in b] == [r["video_id"] for r in a] and len(b) == len(durs)


@pytest.mark.gpu
def test_two_rank_inference_equals_one_rank(tmp_path):
    """`torchrun --nproc-per-node 2 inference.py cfg sub ckpt` writes the same records as one process
    (rank-strided shards, one NCCL all-gather of the kernel-written records). Needs two GPUs."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    durs = [4.03, 5.5, 7.42, 9.04, 6.2, 11.0, 4.7, 8.8, 5.1]
    root, folders, cfg_path, full, sd = _cli_setup(tmp_path, durs)
    outs = []
    for tag, launcher in (("one", [sys.executable]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29611"])):
        ckpt = str(tmp_path / tag / "epoch_010.pth.tar")
        os.makedirs(os.path.dirname(ckpt))
        torch.save({"epoch": 10, "state_dict_ema": {"module." + k: v for k, v in sd.items()}}, ckpt)
        r = subprocess.run(launcher + [os.path.join(root, "inference.py"), cfg_path, "3", ckpt, "-b", "4", "--merge"],
                           capture_output=True, text=True, cwd=root, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(json.load(open(tmp_path / tag / "3" / "data_left.json")))
    assert outs[0] == outs[1]
