"""Process-global configuration, mirror of sres/base/util/config.py:21-106 without hydra.

`cfg()` returns the active configuration (attribute + mapping access like a DictConfig);
`ConfigContext(name, **kwargs)` composes config/<name>.yaml + one YAML per group
(platform, task, model, dataset, pipeline) + dotted overrides such as {'task.nepochs': 100}.
Only one context may be active per process, like the reference (config.py:41,77).
"""
import os
import traceback
from typing import Any, Dict, Optional

import yaml

GROUPS = ("platform", "task", "model", "dataset", "pipeline")


class Cfg(dict):
    """dict with attribute access; nested dicts are wrapped on the way out."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if isinstance(v, dict) and not isinstance(v, Cfg):
            v = Cfg(v)
            dict.__setitem__(self, k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default


def cfg() -> Cfg:
    return ConfigContext.cfg


def config() -> Dict:
    return ConfigContext.configuration


def cid() -> str:
    return "-".join([cfg().model.name, cfg().task.dataset, cfg().task.name])


def cfgdir() -> str:
    return os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "config"))


class ConfigContext:
    cfg: Optional[Cfg] = None
    defaults: Dict = {}
    configuration: Dict = {}

    def __init__(self, name: str, **kwargs):
        assert ConfigContext.cfg is None, "Only one ConfigContext instance is allowed at a time"
        self.name = name
        ConfigContext.configuration = dict(**self.defaults, **kwargs)
        self.model: str = self.get_config("model")
        self.pipeline: str = self.get_config("pipeline", "sres")
        self.platform: str = self.get_config("platform", "local")
        self.task: str = self.get_config("task")
        self.dataset: str = self.get_config("dataset")
        self.config_path: str = self.get_config("config_path", cfgdir())
        self.cid = "-".join([str(s) for s in (self.name, self.model, self.dataset, self.task)])

    def get_config(self, name: str, default: Any = None):
        return self.configuration.get(name, self.defaults.get(name, default))

    @classmethod
    def set_defaults(cls, **kwargs):
        cls.defaults = kwargs

    @classmethod
    def deactivate(cls):
        cls.cfg = None

    @classmethod
    def activate_global(cls, name: str, **kwargs) -> "ConfigContext":
        cc = ConfigContext(name, **kwargs)
        cc.activate()
        return cc

    def _load_group(self, group: str, choice: str) -> dict:
        path = os.path.join(self.config_path, group, f"{choice}.yaml")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"config group '{group}': no such option '{choice}' ({path})")
        with open(path) as f:
            return yaml.safe_load(f) or {}

    def load(self) -> Cfg:
        assert ConfigContext.cfg is None, "Another Config context has already been activated"
        out = Cfg()
        for group in GROUPS:
            choice = getattr(self, group)
            if choice is None:
                raise ValueError(f"ConfigContext: no choice for config group '{group}'")
            out[group] = self._load_group(group, choice)
        for key, val in self.configuration.items():
            if "." in key:  # dotted override, e.g. 'task.nepochs': 100 or 'model.nlayers': 4
                node = out
                parts = key.split(".")
                for p in parts[:-1]:
                    node = node[p]
                node[parts[-1]] = val
        return out

    def activate(self):
        assert ConfigContext.cfg is None, "Context already activated"
        c = ConfigContext.cfg = self.load()
        gpu = self.configuration.get("gpu", int(os.getenv("FMOD_GPU", c.pipeline.get("gpu", 0))))
        c.pipeline.gpu = gpu
        c.task.name = self.task
        c.task.dataset = self.dataset
        c.task.training_version = self.cid

    def __enter__(self):
        self.activate()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.deactivate()
        if exc_type is not None:
            traceback.print_exception(exc_type, value=exc_val, tb=exc_tb)
        return False
