"""systems/utils.py:34-53 of the reference: optimizer by name from torch.optim over model.parameters()."""
import torch


def parse_optimizer(config, model):
    cfg = dict(config)
    params = model.parameters()
    return getattr(torch.optim, cfg["name"])(params, **dict(cfg.get("args", {})))
