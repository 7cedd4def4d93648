from .nms import batched_nms
from .infer_utils import inference_one_epoch, inference_sharded, fix_random_seed, AverageMeter
from .results import merge_results, filter_segments, video_probability

__all__ = ["batched_nms", "inference_one_epoch", "inference_sharded", "fix_random_seed", "AverageMeter", "merge_results", "filter_segments",
           "video_probability"]
