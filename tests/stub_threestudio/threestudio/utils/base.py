"""utils/base.py of the reference, interface only: Updateable (:21-58), BaseObject (:70-88), BaseModule (:91-118)."""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .config import parse_structured
from .misc import get_device, load_module_weights


class Updateable:
    def do_update_step(self, epoch, global_step, on_load_weights=False):
        for name in self.__dir__():
            if name.startswith("_"):
                continue
            try:
                child = getattr(self, name)
            except Exception:
                continue
            if isinstance(child, Updateable):
                child.do_update_step(epoch, global_step, on_load_weights=on_load_weights)
        self.update_step(epoch, global_step, on_load_weights=on_load_weights)

    def do_update_step_end(self, epoch, global_step):
        for name in self.__dir__():
            if name.startswith("_"):
                continue
            try:
                child = getattr(self, name)
            except Exception:
                continue
            if isinstance(child, Updateable):
                child.do_update_step_end(epoch, global_step)
        self.update_step_end(epoch, global_step)

    def update_step(self, epoch, global_step, on_load_weights=False):
        pass

    def update_step_end(self, epoch, global_step):
        pass


class BaseObject(Updateable):
    @dataclass
    class Config:
        pass

    def __init__(self, cfg=None, *args, **kwargs):
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.device = get_device()
        self.configure(*args, **kwargs)

    def configure(self, *args, **kwargs):
        pass


class BaseModule(nn.Module, Updateable):
    @dataclass
    class Config:
        weights: Optional[str] = None

    def __init__(self, cfg=None, *args, **kwargs):
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.device = get_device()
        self.configure(*args, **kwargs)
        if self.cfg.weights is not None:
            path, module_name = self.cfg.weights.split(":")
            sd, epoch, global_step = load_module_weights(path, module_name=module_name, map_location="cpu")
            self.load_state_dict(sd)
            self.do_update_step(epoch, global_step, on_load_weights=True)
        self.register_buffer("_dummy", torch.zeros(0).float(), persistent=False)

    def configure(self, *args, **kwargs):
        pass
