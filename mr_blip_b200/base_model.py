"""BaseModel surface of lavis/models/base_model.py:19-118 plus helpers to hold parameters under the
reference's state-dict key names (SURVEY.md §5 "checkpoint" row: key names are part of the drop-in
contract)."""
import logging
import os

import torch
import torch.nn as nn


class ParamNode(nn.Module):
    """A bare container; children/parameters are attached by dotted name."""


def attach(root: nn.Module, name: str, tensor, requires_grad=False, buffer=False):
    """root.<a>.<b>.<leaf> = Parameter(tensor), creating ParamNode children as needed."""
    parts = name.split(".")
    node = root
    for p in parts[:-1]:
        child = node._modules.get(p)
        if child is None:
            child = ParamNode()
            node.add_module(p, child)
        node = child
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        if not isinstance(tensor, nn.Parameter):
            tensor = nn.Parameter(tensor, requires_grad=requires_grad)
        node.register_parameter(parts[-1], tensor)
    return tensor


class BaseModel(nn.Module):
    """lavis/models/base_model.py:19-118."""

    PRETRAINED_MODEL_CONFIG_DICT = {}

    @property
    def device(self):
        return next(self.parameters()).device

    def load_checkpoint(self, url_or_filename):
        """Non-strict load of a (possibly partial) checkpoint {'model': state_dict} (base_model.py:29-56).
        Only local files: the build environment has no network."""
        if not os.path.isfile(url_or_filename):
            raise RuntimeError("checkpoint url or path is invalid: %s" % url_or_filename)
        ckpt = torch.load(url_or_filename, map_location="cpu")
        state = ckpt["model"] if "model" in ckpt else ckpt
        msg = self.load_state_dict(state, strict=False)
        logging.info("Missing keys %s", msg.missing_keys)
        logging.info("load checkpoint from %s", url_or_filename)
        self._weights_changed()
        return msg

    def _weights_changed(self):
        pass

    @classmethod
    def default_config_path(cls, model_type):
        assert model_type in cls.PRETRAINED_MODEL_CONFIG_DICT, "Unknown model type {}".format(model_type)
        here = os.path.dirname(os.path.abspath(__file__))
        return os.path.join(here, cls.PRETRAINED_MODEL_CONFIG_DICT[model_type])

    def show_n_params(self, return_str=True):
        tot = sum(p.numel() for p in self.parameters())
        if return_str:
            return "{:.1f}M".format(tot / 1e6) if tot >= 1e6 else "{:.1f}K".format(tot / 1e3)
        return tot


def disabled_train(self, mode=True):
    """blip2.py:107-110: keeps the frozen ViT in eval mode."""
    return self
