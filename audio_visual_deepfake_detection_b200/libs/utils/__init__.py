from .nms import batched_nms
from .infer_utils import inference_one_epoch, fix_random_seed, AverageMeter

__all__ = ["batched_nms", "inference_one_epoch", "fix_random_seed", "AverageMeter"]
