from .datasets import make_dataset, make_data_loader, make_inference_dataset, register_dataset, trivial_batch_collator
from . import deepfake_video_audio  # registers the inference datasets

__all__ = ["make_dataset", "make_data_loader", "make_inference_dataset", "register_dataset", "trivial_batch_collator"]
