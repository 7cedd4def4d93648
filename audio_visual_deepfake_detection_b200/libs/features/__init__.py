"""Upstream feature extractors. Only the BYOL-A audio extractor is built: it is the one whose model
code is self-contained in the reference tree (audio_feature/content_audio/byol_a/models.py)."""
from .byola import AudioNTT2020Task6, BatchPlan, LogMelSpectrogram, CONFIG, NORM_STATS

__all__ = ["AudioNTT2020Task6", "BatchPlan", "LogMelSpectrogram", "CONFIG", "NORM_STATS"]
