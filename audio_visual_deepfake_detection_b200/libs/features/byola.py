"""BYOL-A content-audio features on the GPU: host-side mirror of the reference's extractor interface.

Reference interface this replaces (audio_feature/content_audio):
  * extract_audio_feature_one.py:34-46   `to_melspec = MelSpectrogram(...)`, `normalizer = PrecomputedNorm(stats)`,
                                          `model = AudioNTT2020Task6(d=cfg.feature_d, n_mels=64)`, `model.load_weight(...)`
  * extract_audio_feature_one.py:60-75   per file: wav -> lms -> `model(lms.unsqueeze(0))` -> `[T, 2048]` .npy
  * byol_a/models.py:20-40, 48-86        `load_weight` key filtering, the network itself
  * config.yaml:4-11                     16 kHz, n_fft = win = 1024, hop 160, 64 mels, 60-7800 Hz, feature_d 2048

Everything numeric runs in libavdf_sm100.so (no CPU fallback): `avdf_logmel` (FFT + mel + log + normalise),
`avdf_byola_conv1_pool`, `avdf_conv_gemm` with a nine-entry tap table for the two 64->64 3x3 convolutions (tcgen05 path;
BatchNorm folded into weights / bias, ReLU in the epilogue), `avdf_byola_pool`, and `avdf_conv_gemm` again for the two fc
layers. A batch of clips of different lengths is packed along time (csrc/byola.cu "grid layout"): each clip sees exactly
the zero padding it gets alone at batch size 1, which is how the reference script runs it.
"""
import ctypes
import math
import os
import re

import numpy as np
import torch

from ... import native, ops

CONFIG = dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160, n_mels=64, f_min=60.0, f_max=7800.0, feature_d=2048)
NORM_STATS = (-2.2800865, 3.5897882)       # extract_audio_feature_one.py:31
_DT = {"fp32": torch.float32, "mixed": torch.float16, "bf16": torch.bfloat16}


def _pad_to(n, m):
    return (n + m - 1) // m * m


def mel_filterbank(n_freqs=513, f_min=60.0, f_max=7800.0, n_mels=64, sample_rate=16000):
    """The HTK-scale triangular filters torchaudio's MelSpectrogram builds by default (norm = None): [n_freqs, n_mels].
    One-time table set-up, evaluated with torch's fp32 element-wise ops so that the table equals torchaudio's bit for bit."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_lo, m_hi = 2595.0 * math.log10(1.0 + f_min / 700.0), 2595.0 * math.log10(1.0 + f_max / 700.0)
    edges = 700.0 * (10.0 ** (torch.linspace(m_lo, m_hi, n_mels + 2) / 2595.0) - 1.0)
    width = edges[1:] - edges[:-1]
    dist = edges.unsqueeze(0) - freqs.unsqueeze(1)
    falling = (-1.0 * dist[:, :-2]) / width[:-1]
    rising = dist[:, 2:] / width[1:]
    return torch.clamp(torch.minimum(falling, rising), min=0.0).numpy().astype(np.float32)


class BatchPlan:
    """Index arrays of one batch of clips (sample counts `n_samples`): where every clip's frames / pooled time steps live in
    the packed buffers of csrc/byola.cu. Built on the host with numpy, uploaded as two small tensors."""

    def __init__(self, n_samples, device, frames=None):
        if frames is None:
            n_samples = [int(n) for n in n_samples]
            if not n_samples or min(n_samples) <= CONFIG["n_fft"] // 2:
                raise ValueError("every clip needs more than n_fft / 2 = 512 samples (reflect padding)")
            frames = [1 + n // CONFIG["hop_length"] for n in n_samples]          # centre = True
        else:                             # spectrograms supplied by the caller: no waveform behind the plan
            frames = [int(f) for f in frames]
            n_samples = [0] * len(frames)
        if not frames or min(frames) < 8:
            raise ValueError("every clip needs at least 8 spectrogram frames (three 2x2 poolings; torch's max_pool2d refuses fewer too)")
        self.n_clips = len(frames)
        self.frames = frames
        self.t = [[f >> l for f in self.frames] for l in range(4)]          # time steps of a clip at pooling level l

        def grid(tl):                 # one zero step in front of every clip and behind the last; steps padded to 64
            steps = _pad_to(1 + sum(t + 1 for t in tl), 64)
            clip = np.full(steps, -1, np.int32); tt = np.zeros(steps, np.int32); start = np.zeros(len(tl), np.int32)
            s = 1
            for c, t in enumerate(tl):
                start[c] = s
                clip[s:s + t] = c; tt[s:s + t] = np.arange(t)
                s += t + 1
            return steps, clip, tt, start

        self.sample_off = np.concatenate([[0], np.cumsum(n_samples)]).astype(np.int64)
        self.frame_off = np.concatenate([[0], np.cumsum(self.frames)]).astype(np.int32)
        self.total_frames = int(self.frame_off[-1])
        pair_off = np.concatenate([[0], np.cumsum([(f + 1) // 2 for f in self.frames])]).astype(np.int32)     # the FFT kernel takes two frames of a clip at a time
        self.total_pairs = int(pair_off[-1])
        self.steps1, c1, t1, s1 = grid(self.t[1])
        self.steps2, c2, t2, s2 = grid(self.t[2])
        rows3 = sum(self.t[3])
        self.rows3 = _pad_to(max(rows3, 1), 128)
        c3 = np.full(self.rows3, -1, np.int32); t3 = np.zeros(self.rows3, np.int32)
        self.row_off3 = np.concatenate([[0], np.cumsum(self.t[3])]).astype(np.int64)
        for c, t in enumerate(self.t[3]):
            c3[self.row_off3[c]:self.row_off3[c + 1]] = c; t3[self.row_off3[c]:self.row_off3[c + 1]] = np.arange(t)
        parts = {"frame_off": self.frame_off, "pair_off": pair_off, "clip1": c1, "t1": t1, "start1": s1, "clip2": c2, "t2": t2, "start2": s2, "clip3": c3, "t3": t3}
        dev = torch.from_numpy(np.concatenate(list(parts.values()))).to(device)
        o = 0
        for k, v in parts.items():
            setattr(self, "d_" + k, dev[o:o + v.size]); o += v.size
        self.d_sample_off = torch.from_numpy(self.sample_off).to(device)


class LogMelSpectrogram:
    """`normalizer((to_melspec(wav) + eps).log())` of extract_audio_feature_one.py:66 as one kernel."""

    def __init__(self, device, stats=NORM_STATS):
        self.device = torch.device(device)
        self.stats = (float(stats[0]), float(stats[1]))
        n = CONFIG["n_fft"]
        k = np.arange(n, dtype=np.float64)
        self.window = torch.from_numpy((0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)).to(self.device)       # periodic Hann
        j = np.arange(n, dtype=np.float64)
        tw = np.stack([np.cos(2.0 * np.pi * j / n), -np.sin(2.0 * np.pi * j / n)], 1).astype(np.float32)
        self.twiddle = torch.from_numpy(tw).to(self.device)
        fb = mel_filterbank()             # sparse form: every triangle is one run of consecutive bins
        nz = fb > 0
        lo = nz.argmax(0).astype(np.int32); cnt = nz.sum(0).astype(np.int32)
        w = np.zeros((fb.shape[1], int(cnt.max())), np.float32)
        for m in range(fb.shape[1]):
            assert nz[lo[m]:lo[m] + cnt[m], m].all()
            w[m, :cnt[m]] = fb[lo[m]:lo[m] + cnt[m], m]
        self.mel = tuple(torch.from_numpy(a).to(self.device) for a in (lo, cnt, w))

    def packed(self, wav, plan):
        """wav: the batch's clips back to back (CUDA fp32) -> lms [total_frames, 64] (time-major rows)."""
        lms = torch.empty((plan.total_frames, CONFIG["n_mels"]), dtype=torch.float32, device=self.device)
        ops.logmel(wav, plan.d_sample_off, plan.d_frame_off, plan.d_pair_off, plan.total_pairs, self.window, self.twiddle, self.mel, lms, mean=self.stats[0], std=self.stats[1])
        return lms

    def __call__(self, wavs):
        """list of 1-D fp32 arrays / tensors -> list of [64, frames] CUDA tensors (the reference's lms, without the batch dim)."""
        wavs = [torch.as_tensor(np.asarray(w) if not torch.is_tensor(w) else w, dtype=torch.float32).reshape(-1) for w in wavs]
        plan = BatchPlan([w.numel() for w in wavs], self.device)
        lms = self.packed(torch.cat(wavs).to(self.device), plan)
        return [lms[plan.frame_off[c]:plan.frame_off[c + 1]].t() for c in range(plan.n_clips)]


class AudioNTT2020Task6:
    """Drop-in for byol_a/models.py:48-86 on the inference path (eval mode only: BatchNorm uses its running statistics, the
    Dropout between the fc layers is the identity). `precision`: 'mixed' (fp16 operands, fp32 accumulate; default), 'bf16',
    or 'fp32' (CUDA-core parity mode)."""

    def __init__(self, n_mels=64, d=2048, precision="mixed"):
        if n_mels != CONFIG["n_mels"]:
            raise ValueError("the kernels are built for n_mels = 64 (config.yaml:9)")
        if precision not in _DT:
            raise ValueError("precision must be one of %s" % sorted(_DT))
        self.n_mels, self.d, self.precision = n_mels, d, precision
        self.device = None
        self._sd = None
        self._w = None
        self._melspec = None
        self._stage = None                # grow-only pinned staging buffer of extract()

    # ---- nn.Module-shaped surface the extraction scripts use
    def load_state_dict(self, sd):
        need = ["features.%d.%s" % (i, k) for i in (0, 4, 8) for k in ("weight", "bias")]
        need += ["features.%d.%s" % (i, k) for i in (1, 5, 9) for k in ("weight", "bias", "running_mean", "running_var")]
        need += ["fc.0.weight", "fc.0.bias", "fc.3.weight", "fc.3.bias"]
        missing = [k for k in need if k not in sd]
        if missing:
            raise KeyError("missing keys in state_dict: %s" % missing)
        self._sd = {k: torch.as_tensor(sd[k]).detach().to(torch.float32).cpu() for k in need}
        if tuple(self._sd["fc.0.weight"].shape) != (self.d, 64 * (self.n_mels // 8)):
            raise ValueError("fc.0.weight has shape %s" % (tuple(self._sd["fc.0.weight"].shape),))
        self._w = None
        return self

    def load_weight(self, weight_file, device, state_dict=None, key_check=True):
        """models.py:20-40: strip everything in front of `features.` / `fc.` from the checkpoint's keys."""
        sd = state_dict or torch.load(weight_file, map_location="cpu")
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if key_check:
            kept = {}
            for k, v in sd.items():
                m = re.search(r"(^fc\.|\.fc\.|^features\.|\.features\.)", k)
                if m is not None:
                    kept[k[m.start():].lstrip(".")] = v
            sd = kept
        self.load_state_dict(sd)
        return self.to(device).eval()

    def to(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AudioNTT2020Task6 runs on CUDA devices only (no CPU fallback)")
        self._w = None
        return self

    def eval(self):
        return self

    def state_dict(self):
        return dict(self._sd)

    # ---- weights in the kernels' formats
    def _weights(self):
        if self._w is not None:
            return self._w
        if self._sd is None or self.device is None:
            raise RuntimeError("load_state_dict(...) and to(device) first")
        sd, adt, dev = self._sd, _DT[self.precision], self.device
        w = {}
        for i, (conv, bn) in enumerate(((0, 1), (4, 5), (8, 9))):
            scale = sd[f"features.{bn}.weight"] / torch.sqrt(sd[f"features.{bn}.running_var"] + 1e-5)
            bias = sd[f"features.{bn}.bias"] + (sd[f"features.{conv}.bias"] - sd[f"features.{bn}.running_mean"]) * scale
            cw = sd[f"features.{conv}.weight"] * scale[:, None, None, None]               # [out, in, mel tap, time tap]
            if i == 0:
                w["w1"] = cw.reshape(64, 9).contiguous().to(dev)
            else:                     # rows [out, tap * 64 + in] with tap = time tap * 3 + mel tap (the order of `tap_rows`)
                w[f"w{i + 1}"] = cw.permute(0, 3, 2, 1).reshape(64, 9 * 64).contiguous().to(dev, adt)
            w[f"b{i + 1}"] = bias.contiguous().to(dev)
        w["fc1"] = sd["fc.0.weight"].contiguous().to(dev, adt); w["fb1"] = sd["fc.0.bias"].contiguous().to(dev)
        w["fc2"] = sd["fc.3.weight"].contiguous().to(dev, adt); w["fb2"] = sd["fc.3.bias"].contiguous().to(dev)
        self._w = w
        return w

    # ---- the network on a packed batch
    def forward_packed(self, lms, plan):
        """lms [total_frames, 64] fp32 (time-major, clips back to back as `plan` lays them out) -> [plan.rows3, d] fp32; clip
        c owns rows plan.row_off3[c] : plan.row_off3[c + 1]."""
        w, adt, dev = self._weights(), _DT[self.precision], self.device
        f32 = adt == torch.float32

        def conv3x3(x, key, steps, mel, mask):
            rows = steps * (mel + 2)
            out = torch.empty((rows, 64), dtype=adt, device=dev)
            taps = [dt * (mel + 2) + dm for dt in (-1, 0, 1) for dm in (-1, 0, 1)]
            ops.conv_gemm(x, w["w" + key], taps=9, batch=1, c_in=64, n_out=64, segs=[(rows, 0, 0)], a_rows=rows, o_rows=rows,
                          bias=w["b" + key], row_mask=mask, act=ops.ACT_RELU, tap_rows=taps,
                          **({"out_f32": out} if f32 else {"out_h": out}))
            return out

        g1 = torch.empty((plan.steps1 * 34, 64), dtype=adt, device=dev)
        m1 = torch.empty((plan.steps1 * 34,), dtype=torch.uint8, device=dev)
        ops.byola_conv1_pool(lms, plan.d_frame_off, w["w1"], w["b1"], plan.d_clip1, plan.d_t1, g1, m1)
        c2 = conv3x3(g1, "2", plan.steps1, 32, m1)
        g2 = torch.empty((plan.steps2 * 18, 64), dtype=adt, device=dev)
        m2 = torch.empty((plan.steps2 * 18,), dtype=torch.uint8, device=dev)
        ops.byola_pool(c2, plan.d_start1, plan.d_clip2, plan.d_t2, g2, mel_in=32, pad_out=1, mask_out=m2)
        c3 = conv3x3(g2, "3", plan.steps2, 16, m2)
        x3 = torch.empty((plan.rows3, 512), dtype=adt, device=dev)
        ops.byola_pool(c3, plan.d_start2, plan.d_clip3, plan.d_t3, x3, mel_in=16, pad_out=0)
        h = torch.empty((plan.rows3, self.d), dtype=adt, device=dev)
        kw = dict(taps=1, batch=1, segs=[(plan.rows3, 0, 0)], a_rows=plan.rows3, o_rows=plan.rows3, act=ops.ACT_RELU)
        ops.conv_gemm(x3, w["fc1"], c_in=512, n_out=self.d, bias=w["fb1"], **({"out_f32": h} if f32 else {"out_h": h}), **kw)
        out = torch.empty((plan.rows3, self.d), dtype=torch.float32, device=dev)
        ops.conv_gemm(h, w["fc2"], c_in=self.d, n_out=self.d, bias=w["fb2"], out_f32=out, **kw)
        return out

    def _split(self, out, plan):
        return [out[plan.row_off3[c]:plan.row_off3[c + 1]] for c in range(plan.n_clips)]

    def forward(self, x):
        """x: (B, 1, 64, T) like the reference (all clips of one length) -> (B, T // 8, d); or a list of (64, T_i) /
        (1, 64, T_i) tensors -> list of (T_i // 8, d)."""
        as_list = isinstance(x, (list, tuple))
        clips = [c.reshape(self.n_mels, -1) for c in (x if as_list else x.reshape(x.shape[0], self.n_mels, x.shape[-1]))]
        frames = [int(c.shape[1]) for c in clips]
        plan = BatchPlan(None, self.device, frames=frames)
        lms = torch.cat([c.to(self.device, torch.float32).t() for c in clips]).contiguous()
        outs = self._split(self.forward_packed(lms, plan), plan)
        return list(outs) if as_list else torch.stack(outs)

    __call__ = forward

    def extract(self, wavs, stats=NORM_STATS):
        """The loop body of extract_audio_feature_one.py:60-75 for a batch of clips: list of 1-D 16 kHz fp32 arrays (host
        or device) -> list of [frames // 8, d] fp32 CUDA tensors (what the script saves as `.npy`)."""
        if self._melspec is None or self._melspec.device != self.device or self._melspec.stats != tuple(map(float, stats)):
            self._melspec = LogMelSpectrogram(self.device, stats)
        on_device = all(torch.is_tensor(w) and w.is_cuda for w in wavs)
        if not on_device:
            wavs = [np.ascontiguousarray(w.detach().cpu().numpy() if torch.is_tensor(w) else w, dtype=np.float32).reshape(-1) for w in wavs]
        plan = BatchPlan([(w.numel() if on_device else w.shape[0]) for w in wavs], self.device)
        if on_device:
            wav = torch.cat([w.reshape(-1).float() for w in wavs])
        else:
            n = int(plan.sample_off[-1])
            if self._stage is None or self._stage[0].numel() < n:
                self._stage = (torch.empty(max(2 * n, 1 << 22), dtype=torch.float32).pin_memory(), torch.cuda.Event())
            host, done = self._stage
            done.synchronize()            # the previous batch's copy has left the buffer
            # gather into the pinned buffer with the library's host thread pool (non-temporal stores, csrc/host_pack.cu)
            k = len(wavs)
            base = host.data_ptr()
            src = (ctypes.c_void_p * k)(*[w.ctypes.data for w in wavs])
            dst = (ctypes.c_void_p * k)(*[base + 4 * int(o) for o in plan.sample_off[:-1]])
            nb = (ctypes.c_size_t * k)(*[w.nbytes for w in wavs])
            native.check(native.lib().avdf_host_pack(src, dst, nb, k, min(8, os.cpu_count() or 1)), "avdf_host_pack")
            wav = host[:n].to(self.device, non_blocking=True)
            done.record(torch.cuda.current_stream(self.device))
        lms = self._melspec.packed(wav, plan)
        return self._split(self.forward_packed(lms, plan), plan)
