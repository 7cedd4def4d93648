"""Inference datasets of the hot path: `.npy` feature ingestion.

`deepfake_video_audioEmoBYOLA_inference` (libs/datasets/deepfake_video_audio.py:351-558) and the audio-only variant:
same constructor keywords, same test-list format (`deepfake_test_sub{sub_index}.txt`, lines `id.mp4,duration`), same
truncation of the audio streams (:482-483) and the same `fps / duration / feat_stride / feat_num_frames` metadata
(:461, :495-500).

Difference by design: `__getitem__` does NOT resample on the CPU. It returns the raw time-major streams
(`item['streams']`), and the fixed-length linear interpolation + concat (:513-547) runs as the first CUDA kernel of the
model (`avdf_interp_concat`): `model(video_list)` / `model.stream(batches)` accept these items directly. Callers that
want the reference's `feats [C, T]` tensor call `materialize(item)`, which computes it with the same kernel.
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from .datasets import register_dataset

BYOLA_FPS = 12.497      # deepfake_video_audio.py:415
EMOTION_FPS = 50        # deepfake_video_audio.py:416


def _load_npy_pinned(path):
    """A C-ordered fp32 `.npy` file read directly into a page-locked buffer (file -> pinned memory is the only host copy);
    None for any other layout (the caller falls back to np.load)."""
    with open(path, "rb") as f:
        try:
            major, _ = np.lib.format.read_magic(f)
            shape, fortran, dtype = (np.lib.format.read_array_header_1_0 if major == 1 else np.lib.format.read_array_header_2_0)(f)
        except Exception:
            return None
        if fortran or dtype != np.dtype("<f4") or len(shape) != 2:
            return None
        buf = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
        want = buf.nbytes
        got = f.readinto(memoryview(buf).cast("B")) if want else 0
        if got != want:
            raise IOError("%s: truncated .npy payload (%d of %d bytes)" % (path, got, want))
        return buf


class _InferenceBase(Dataset):
    USE_VIDEO = True

    def __init__(self, is_training, split, sub_index, crop_ratio=None, default_fps=None, downsample_rate=1,
                 video_feat_folder=None, audio_feat_folder=None, audio_byola_feat_folder=None, audio_emo_feat_folder=None,
                 audio_file_ext=None, num_classes=1, input_dim=None, video_input_dim=None, audio_input_dim=None,
                 feat_stride=1, num_frames=1, test_folder=None, trunc_thresh=0.5, max_seq_len=768, force_upsampling=True,
                 pin_memory=False, **unused):
        assert not is_training, "training datasets are outside the accelerated inference path"
        assert num_classes == 1
        if not force_upsampling or feat_stride <= 0:
            raise RuntimeError("not implemented")       # same as the reference's case 3 (:502-503)
        self.sub_index = sub_index
        # new: read the .npy payloads straight into page-locked memory, so that model.stream() hands them to the copy engine
        # without a staging copy (libs/modeling/streaming.py). In-process loading only: pinned pages do not survive the
        # pickling of DataLoader worker processes.
        self.pin_memory = bool(pin_memory)
        self.video_feat_folder = video_feat_folder if self.USE_VIDEO else None
        self.audio_byola_feat_folder = audio_byola_feat_folder
        self.audio_emo_feat_folder = audio_emo_feat_folder
        self.test_folder = test_folder
        self.feat_stride, self.num_frames = feat_stride, num_frames
        self.max_seq_len = max_seq_len
        self.num_classes = num_classes
        self.label_dict = {"Fake": 0}
        self.byola_fps, self.emotion_fps = BYOLA_FPS, EMOTION_FPS
        self.data_list = self._get_test_infos()
        self.db_attributes = {"dataset_name": "DeepFake_Audio", "tiou_thresholds": np.linspace(0.5, 0.95, 10),
                              "empty_label_ids": []}

    def _get_test_infos(self):
        """:420-431"""
        path = os.path.join(self.test_folder, f"deepfake_test_sub{self.sub_index}.txt")
        out = []
        with open(path, "r") as f:
            for line in f:
                items = line.strip().split(",")
                if len(items) >= 2:
                    out.append({"id": items[0], "duration": float(items[1])})
        return out

    def get_attributes(self):
        return self.db_attributes

    def __len__(self):
        return len(self.data_list)

    def _load(self, folder, vid):
        path = os.path.join(folder, vid.replace(".mp4", ".npy"))
        if self.pin_memory:
            arr = _load_npy_pinned(path)
            if arr is not None:
                return arr
        return np.load(path).astype(np.float32, copy=False)

    def __getitem__(self, idx):
        v = self.data_list[idx]
        dur = v["duration"]
        streams = {}
        if self.video_feat_folder is not None:
            streams["video"] = self._load(self.video_feat_folder, v["id"])
        b = self._load(self.audio_byola_feat_folder, v["id"])
        e = self._load(self.audio_emo_feat_folder, v["id"])
        streams["byola"] = b[: int(self.byola_fps * dur - 0.3657)]          # :482
        streams["emo"] = e[: int(self.emotion_fps * dur - 0.817)]           # :483
        first = streams["video"] if "video" in streams else streams["byola"]
        fps = first.shape[0] / dur                                            # :461
        feat_stride = float((first.shape[0] - 1) * self.feat_stride + self.num_frames) / self.max_seq_len   # :495-497
        return {"video_id": v["id"], "streams": streams, "fps": fps, "duration": dur,
                "feat_stride": feat_stride, "feat_num_frames": feat_stride}

    def materialize(self, item, device="cuda"):
        """The reference item: adds `feats` [C, max_seq_len] (CPU fp32), computed by the interp/concat kernel."""
        from ... import ops
        names = [n for n in ("video", "byola", "emo") if n in item["streams"]]
        tens, offs, C = [], [], 0
        for n in ("video", "byola", "emo"):
            if n in item["streams"]:
                a = torch.from_numpy(np.ascontiguousarray(item["streams"][n])).to(device)
                tens.append(a); offs.append(torch.tensor([0, a.shape[0]], dtype=torch.int32, device=device)); C += a.shape[1]
            else:
                tens.append(None); offs.append(None)
        out = torch.empty((1, self.max_seq_len, C), dtype=torch.float32, device=device)
        ops.interp_concat(tens, offs, self.max_seq_len, out)
        full = dict(item)
        full["feats"] = out[0].t().contiguous().cpu()
        del names
        return full


@register_dataset("deepfake_video_audioEmoBYOLA_inference")
class DeepFakeVideoAudioDatasetInfer3(_InferenceBase):
    """visual (256) | BYOL-A (2048) | emotion2vec (768)"""
    USE_VIDEO = True


@register_dataset("deepfake_audioEmoBYOLA_inference")
class DeepFakeAudioEmoByolaDatasetInfer(_InferenceBase):
    """audio-only: BYOL-A (2048) | emotion2vec (768) (BASELINE.json configs[1])"""
    USE_VIDEO = False
