"""Run-config loader with the merge order of lavis/common/config.py:17-112 (model defaults of (arch, model_type) <- the
recipe's `model` section <- `--options key=value` overrides), on plain dicts: the reference builds it on OmegaConf, which
this environment does not have ("loader shim").  Recipes written for the reference
(lavis/projects/mr_BLIP/**.yaml) load unchanged; `run` and `datasets` sections are passed through for the runner.

    cfg = Config("mr_blip_b200/configs/projects/mr_BLIP/train/qvh.yaml", options=["model.input_time_format=seconds_floats"])
    model = registry.get_model_class(cfg.model_cfg["arch"]).from_config(cfg.model_cfg)
"""
import re

import yaml

from .registry import registry


class _Loader(yaml.SafeLoader):
    """SafeLoader whose floats include the exponent-without-dot form (`init_lr: 3e-4` in every mr_BLIP recipe), which
    YAML 1.1 / plain PyYAML reads as a string and OmegaConf reads as a float."""


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9_]+(?:[eE][-+][0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)
                    |\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."))


def _load_yaml(text):
    return yaml.load(text, Loader=_Loader)


class Section(dict):
    """dict with attribute access and the .get(key, default) the from_config methods use."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _wrap(x):
    if isinstance(x, dict):
        return Section({k: _wrap(v) for k, v in x.items()})
    return x


def _deep_update(dst, src):
    """Nested merge: --options datasets.qvh.build_info.videos.storage=... must not drop the section's other keys."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        else:
            dst[k] = v
    return dst


def _parse_value(s):
    return _load_yaml(s)


class Config:
    def __init__(self, cfg_path, options=None):
        with open(cfg_path) as f:
            user = _load_yaml(f.read()) or {}
        from . import blip2_mr, blip2_t5  # noqa: F401  (register the model classes)
        opts = self._options(options)
        user_model = dict(user.get("model", {}))
        user_model.update(opts.get("model", {}))
        assert "arch" in user_model, "Missing model configuration file."                  # config.py:60
        model_cls = registry.get_model_class(user_model["arch"])
        assert model_cls is not None, "Model '%s' has not been registered." % user_model["arch"]
        model_type = user_model.get("model_type", None)
        assert model_type is not None, "Missing model_type."                               # config.py:68
        with open(model_cls.default_config_path(model_type)) as f:
            defaults = (_load_yaml(f.read()) or {}).get("model", {})
        merged = dict(defaults)
        merged.update(user_model)                                                          # user overrides the defaults
        self.model_cfg = _wrap(merged)
        run = dict(user.get("run", {}))
        run.update(opts.get("run", {}))
        self.run_cfg = _wrap(run)
        ds = {k: dict(v or {}) for k, v in (user.get("datasets", {}) or {}).items()}
        _deep_update(ds, opts.get("datasets", {}))
        self.datasets_cfg = _wrap(ds)

    @staticmethod
    def _options(options):
        """["model.task=qformer_freeze_lora", "run.batch_size_train=4"] -> nested dict (config.py:83-112 dotlist)."""
        out = {}
        for item in options or []:
            key, _, val = item.partition("=")
            node = out
            parts = key.strip().split(".")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = _parse_value(val)
        return out

    def to_dict(self):
        """Plain nested dicts {"run", "model", "datasets"} -- what the runner stores in a checkpoint (config.py:165-166)."""
        def plain(x):
            return {k: plain(v) for k, v in x.items()} if isinstance(x, dict) else x
        return {"run": plain(self.run_cfg), "model": plain(self.model_cfg), "datasets": plain(self.datasets_cfg)}

    def n_frames(self, split="train"):
        """Frames per clip the video processors of the (single) dataset sample (datasets.*.vis_processor.<split>.n_frms)."""
        for d in self.datasets_cfg.values():
            return d.get("vis_processor", {}).get(split, {}).get("n_frms")
        return None


def build_model(cfg_path, options=None, **overrides):
    """Recipe -> model instance (what lavis/tasks/base_task.py:34-37 build_model does with its Config)."""
    cfg = Config(cfg_path, options)
    cfg.model_cfg.update(overrides)
    cls = registry.get_model_class(cfg.model_cfg["arch"])
    return cls.from_config(cfg.model_cfg), cfg
