"""Load the reference's own hot-path modules (eva_vit.py, Qformer.py, modeling_t5.py) by file
path from /root/reference, with the minimal shims SURVEY.md §8c lists.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to generate the committed golden
vectors) and by tests that are skipped when /root/reference is absent (the GPU box).  Nothing in
the product (mr_blip_b200/) or in bench.py imports this file.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MRB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lavis", "models"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package so sub-imports resolve
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _install_shims():
    import torch
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    # timm (eva_vit.py:15-16)
    if "timm" not in sys.modules:
        def drop_path(x, drop_prob=0.0, training=False):
            if drop_prob == 0.0 or not training:
                return x
            keep = 1 - drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = x.new_empty(shape).bernoulli_(keep)
            return x * mask / keep

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", drop_path=drop_path, to_2tuple=to_2tuple, trunc_normal_=trunc_normal_)
        _stub("timm.models.registry", register_model=lambda f: f)
    # lavis.common.dist_utils.download_cached_file (eva_vit.py:18)
    if "lavis" not in sys.modules:
        _stub("lavis")
        _stub("lavis.common")
        _stub("lavis.common.dist_utils", download_cached_file=lambda *a, **k: (_ for _ in ()).throw(
            RuntimeError("offline: no download")))
    # transformers symbols that moved/vanished after 4.46.1 (Qformer.py:39-44, modeling_t5.py:37-51)
    def find_pruneable_heads_and_indices(heads, n_heads, head_size, already_pruned_heads):
        raise NotImplementedError("head pruning is not on the hot path")

    for mod in (mu, pu):
        if not hasattr(mod, "apply_chunking_to_forward"):
            mod.apply_chunking_to_forward = pu.apply_chunking_to_forward
        if not hasattr(mod, "prune_linear_layer"):
            mod.prune_linear_layer = pu.prune_linear_layer
        if not hasattr(mod, "find_pruneable_heads_and_indices"):
            mod.find_pruneable_heads_and_indices = find_pruneable_heads_and_indices
    try:
        import transformers.utils.model_parallel_utils  # noqa: F401
    except Exception:
        _stub("transformers.utils.model_parallel_utils",
              assert_device_map=lambda *a, **k: None, get_device_map=lambda *a, **k: None)
    import transformers.utils as tu
    for name in ("DUMMY_INPUTS", "DUMMY_MASK"):
        if not hasattr(tu, name):
            setattr(tu, name, [[0]])
    for name in ("add_start_docstrings", "add_start_docstrings_to_model_forward"):
        if not hasattr(tu, name):
            setattr(tu, name, lambda *a, **k: (lambda f: f))
    if not hasattr(tu, "replace_return_docstrings"):
        tu.replace_return_docstrings = lambda *a, **k: (lambda f: f)
    if not hasattr(tu, "is_torch_fx_proxy"):
        tu.is_torch_fx_proxy = lambda x: False
    import transformers.file_utils as fu
    if not hasattr(fu, "ModelOutput"):
        fu.ModelOutput = tu.ModelOutput
    PT = mu.PreTrainedModel
    if not getattr(PT, "_mrb_shimmed", False):
        PT.get_head_mask = lambda self, m, n, *a, **k: [None] * n
        _orig_init_weights = PT.init_weights

        def init_weights(self):
            try:
                return _orig_init_weights(self)
            except Exception:
                return self.apply(self._init_weights)

        PT.init_weights = init_weights
        PT._mrb_shimmed = True


def _load(name, relpath):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    """-> (eva_vit, Qformer, modeling_t5) reference modules, executed unmodified."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    eva = _load("_ref_eva_vit", "lavis/models/eva_vit.py")
    qf = _load("_ref_Qformer", "lavis/models/blip2_models/Qformer.py")
    t5 = _load("_ref_modeling_t5", "lavis/models/blip2_models/modeling_t5.py")
    return eva, qf, t5


def load_reference_utils():
    """reference blip2_mr_models/utils.py (post_process etc.); needs av/wandb stubs."""
    _install_shims()
    for n in ("av", "wandb"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n, run=None)
    return _load("_ref_mr_utils", "lavis/models/blip2_mr_models/utils.py")
