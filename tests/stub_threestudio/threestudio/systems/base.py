"""systems/base.py of the reference, interface only, without Lightning: BaseSystem (:21-200: cfg parsing, configure(),
C(), configure_optimizers(), the per-batch do_update_step walk) and BaseLift3DSystem (:213-300: geometry / material /
background / renderer built through the registry; renderer gets the three modules as keyword arguments)."""
from dataclasses import dataclass, field
from typing import List, Optional

import torch.nn as nn

import threestudio
from threestudio.utils.base import Updateable
from threestudio.utils.config import parse_structured
from threestudio.utils.misc import C
from .utils import parse_optimizer


class BaseSystem(nn.Module, Updateable):
    @dataclass
    class Config:
        loggers: dict = field(default_factory=dict)
        loss: dict = field(default_factory=dict)
        optimizer: dict = field(default_factory=dict)
        scheduler: Optional[dict] = None
        weights: Optional[str] = None
        weights_ignore_modules: Optional[List[str]] = None
        cleanup_after_validation_step: bool = False
        cleanup_after_test_step: bool = False

    def __init__(self, cfg, resumed=False):
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self._save_dir = None
        self._resumed = resumed
        self.global_step = 0          # pl.LightningModule properties in the reference
        self.current_epoch = 0
        self.logged = {}
        self.configure()
        self.post_configure()

    @property
    def resumed(self):
        return self._resumed

    @property
    def true_global_step(self):
        return self.global_step

    @property
    def true_current_epoch(self):
        return self.current_epoch

    def configure(self):
        pass

    def post_configure(self):
        pass

    def C(self, value):
        return C(value, self.true_current_epoch, self.true_global_step)

    def log(self, name, value, **kw):
        self.logged[name] = value

    def configure_optimizers(self):
        return {"optimizer": parse_optimizer(self.cfg.optimizer, self)}

    def on_train_batch_start(self, batch, batch_idx, unused=0):
        self.do_update_step(self.true_current_epoch, self.true_global_step)

    def on_train_batch_end(self, outputs, batch, batch_idx):
        self.do_update_step_end(self.true_current_epoch, self.true_global_step)

    def on_fit_start(self):
        pass


class BaseLift3DSystem(BaseSystem):
    @dataclass
    class Config(BaseSystem.Config):
        geometry_type: str = ""
        geometry: dict = field(default_factory=dict)
        geometry_convert_from: Optional[str] = None
        geometry_convert_inherit_texture: bool = False
        geometry_convert_override: dict = field(default_factory=dict)
        material_type: str = ""
        material: dict = field(default_factory=dict)
        background_type: str = ""
        background: dict = field(default_factory=dict)
        renderer_type: str = ""
        renderer: dict = field(default_factory=dict)
        guidance_type: str = ""
        guidance: dict = field(default_factory=dict)
        prompt_processor_type: str = ""
        prompt_processor: dict = field(default_factory=dict)
        exporter_type: str = "mesh-exporter"
        exporter: dict = field(default_factory=dict)

    def configure(self):
        self.geometry = threestudio.find(self.cfg.geometry_type)(self.cfg.geometry)
        self.material = threestudio.find(self.cfg.material_type)(self.cfg.material)
        self.background = threestudio.find(self.cfg.background_type)(self.cfg.background)
        self.renderer = threestudio.find(self.cfg.renderer_type)(self.cfg.renderer, geometry=self.geometry,
                                                                 material=self.material, background=self.background)
