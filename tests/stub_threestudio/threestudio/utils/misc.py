"""utils/misc.py of the reference: get_device (:28-29), C (:65-86), cleanup, load_module_weights -- interface only."""
import gc
import os

import torch


def get_rank():
    for k in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "JSM_NAMESPACE_RANK"):
        if k in os.environ:
            return int(os.environ[k])
    return 0


def get_device():
    # the reference hard-codes cuda:{rank}; the stub falls back to the CPU so construction-only tests run without a GPU
    return torch.device(f"cuda:{get_rank()}") if torch.cuda.is_available() else torch.device("cpu")


def C(value, epoch, global_step):
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    start, v0, v1, end = value
    cur = global_step if isinstance(end, int) else epoch
    return v0 + (v1 - v0) * max(min(1.0, (cur - start) / (end - start)), 0.0)


def cleanup():
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def load_module_weights(path, module_name=None, ignore_modules=None, map_location=None):
    ckpt = torch.load(path, map_location=map_location)
    sd = ckpt["state_dict"]
    if module_name is not None:
        sd = {k[len(module_name) + 1:]: v for k, v in sd.items() if k.startswith(module_name + ".")}
    return sd, ckpt.get("epoch", 0), ckpt.get("global_step", 0)
