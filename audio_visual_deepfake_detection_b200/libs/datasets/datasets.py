"""Dataset registry and loader builders with the reference's surface (libs/datasets/datasets.py:5-43)."""
import torch

datasets = {}


def register_dataset(name):
    def decorator(cls):
        datasets[name] = cls
        return cls
    return decorator


def trivial_batch_collator(batch):
    """libs/datasets/data_utils.py:9-13: the model does its own batching."""
    return batch


def make_dataset(name, is_training, split, **kwargs):
    return datasets[name](is_training, split, **kwargs)


def make_inference_dataset(name, is_training, split, sub_index, **kwargs):
    return datasets[name](is_training, split, sub_index, **kwargs)


def make_data_loader(dataset, is_training, generator, batch_size, num_workers):
    """Same arguments as the reference (datasets.py:28-43). batch_size may be > 1 here: the accelerated model takes any
    number of videos per call (the reference asserts 1, av_fd_no_recon.py:456)."""
    return torch.utils.data.DataLoader(
        dataset, batch_size=batch_size, num_workers=num_workers, collate_fn=trivial_batch_collator, shuffle=is_training,
        drop_last=is_training, generator=generator, persistent_workers=num_workers > 0)
