"""Inference driver with the reference's interface (libs/utils/train_utils.py:510-596
`inference_one_epoch`, :22-40 `fix_random_seed`). Training utilities (optimizer, scheduler,
EMA, checkpoint saving) are outside the accelerated path."""
import json
import os
import random

import numpy as np
import torch


class AverageMeter(object):
    def __init__(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def fix_random_seed(seed, include_cuda=True):
    gen = torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    if include_cuda and torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return gen


def result_item(video_id, out):
    """The JSON record of train_utils.py:577-591."""
    n = out["segments"].shape[0]
    return {"video_id": video_id, "video_cls": out["video_cls"].numpy().tolist(),
            "scores": out["scores"].numpy().tolist() if n else [],
            "segments": out["segments"].numpy().tolist() if n else []}


def inference_one_epoch(val_loader, model, curr_epoch, ext_score_file=None, evaluator=None, output_folder=None,
                        tb_writer=None, print_freq=20, subset="test", max_avg_nr_proposal=100, dataset_name="",
                        dump_every=5000):
    """Runs `model(video_list)` over the loader and writes `data_left<iter>.json` every `dump_every`
    iterations plus a final `data_left.json` (train_utils.py:546-595). Returns the records of the last dump
    window (the reference returns None; callers ignore it)."""
    assert (evaluator is not None) or (output_folder is not None)
    model.eval()
    os.makedirs(output_folder, exist_ok=True)
    batch_results = []
    for iter_idx, video_list in enumerate(val_loader, 0):
        if iter_idx > 0 and iter_idx % dump_every == 0:
            with open(f"{output_folder}/data_left{iter_idx}.json", "w", encoding="utf-8") as f:
                json.dump(batch_results, f, ensure_ascii=False, indent=4)
            batch_results = []
        with torch.no_grad():
            output = model(video_list)
        for vid_idx in range(len(output)):
            batch_results.append(result_item(video_list[vid_idx]["video_id"], output[vid_idx]))
    if len(batch_results) > 0:
        with open(f"{output_folder}/data_left.json", "w", encoding="utf-8") as f:
            json.dump(batch_results, f, ensure_ascii=False, indent=4)
    return batch_results
