"""Golden learning-rate traces from the REFERENCE's own schedulers (lavis/common/optims.py, loaded by file path with a
stub registry) for the recipe settings the mr_BLIP projects use and a few edge cases.
Run in the build container (needs /root/reference): python tests/golden/make_golden_lr.py
-> tests/golden/lr_sched_golden.json"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MRB_REFERENCE_ROOT", "/root/reference")

CASES = [  # scheduler, kwargs, epochs, iters per epoch
    ("linear_warmup_cosine_lr", dict(max_epoch=50, min_lr=0.0, init_lr=3e-4, warmup_start_lr=1e-8, warmup_steps=15), 4, 11),
    ("linear_warmup_cosine_lr", dict(max_epoch=20, min_lr=1e-6, init_lr=3e-4, warmup_start_lr=1e-8, warmup_steps=40), 5, 13),
    ("linear_warmup_cosine_lr", dict(max_epoch=3, min_lr=1e-5, init_lr=1e-4, warmup_start_lr=-1, warmup_steps=0), 3, 4),
    ("linear_warmup_step_lr", dict(max_epoch=6, min_lr=1e-5, init_lr=1e-4, decay_rate=0.5, warmup_start_lr=1e-6, warmup_steps=5), 6, 7),
    ("linear_warmup_step_lr", dict(max_epoch=3, min_lr=0.0, init_lr=2e-4, decay_rate=0.9, warmup_start_lr=-1, warmup_steps=0), 3, 3),
]


class _Opt:
    def __init__(self):
        self.param_groups = [{"lr": None}, {"lr": None}]


def trace(cls, kw, epochs, iters):
    opt = _Opt()
    s = cls(optimizer=opt, **kw)
    out = []
    for e in range(epochs):
        for i in range(iters):
            s.step(cur_epoch=e, cur_step=i)
            assert opt.param_groups[0]["lr"] == opt.param_groups[1]["lr"]
            out.append(opt.param_groups[0]["lr"])
    return out


def main():
    table = {}

    class _Reg:
        @staticmethod
        def register_lr_scheduler(name):
            def wrap(c):
                table[name] = c
                return c
            return wrap

    for name in ("lavis", "lavis.common"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    reg = types.ModuleType("lavis.common.registry")
    reg.registry = _Reg
    sys.modules["lavis.common.registry"] = reg
    spec = importlib.util.spec_from_file_location("ref_optims", os.path.join(REF_ROOT, "lavis/common/optims.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = [{"sched": n, "kwargs": kw, "epochs": e, "iters": it, "lr": trace(table[n], kw, e, it)} for n, kw, e, it in CASES]
    json.dump(out, open(os.path.join(HERE, "lr_sched_golden.json"), "w"), indent=0)
    print("wrote", len(out), "traces")


if __name__ == "__main__":
    main()
