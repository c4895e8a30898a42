"""Stand-alone training / evaluation loop for the moment-retrieval recipes, without the LAVIS package:

    python -m mr_blip_b200.train --cfg-path mr_blip_b200/configs/projects/mr_BLIP/train/qvh.yaml \\
        --options datasets.qvh.build_info.annotations.train.storage=/data/qvh/train.json \\
                  datasets.qvh.build_info.annotations.val.storage=/data/qvh/val.json \\
                  datasets.qvh.build_info.videos.storage=/data/qvh/videos run.output_dir=result/qvh
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m mr_blip_b200.train --cfg-path ...      # one process per GPU

It is the control flow of the reference's train.py -> RunnerBase.train (lavis/runners/runner_base.py:362-420) ->
MomentRetrievalTask._train_inner_loop / evaluation / after_evaluation (lavis/tasks/moment_retrieval.py:33-257) reduced to
what the mr_BLIP recipes use: epoch-based, lr stepped every iteration, gradient accumulation, validation after every epoch,
best checkpoint by agg_metrics (mean R1).  Inside LAVIS none of this is needed -- the runner drives the model class directly
(INTEGRATION.md).  Differences: no GradScaler / autocast (the kernels choose their own operand types and hand back an fp32
loss), gradients are averaged across ranks once per optimiser step on the model's flat buffer instead of per micro-step by
DDP (the same sum), frames travel as uint8."""
import argparse
import json
import logging
import os
import random

import numpy as np
import torch

from . import checkpoint, data, dist as mdist, mr_eval, optim
from .config import Config
from .registry import registry


def prepare_sample(samples, device):
    """Tensors of the collated sample dict to `device`, non-blocking (lavis/datasets/data_utils.py:168-174)."""
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in samples.items()}


class MomentRetrievalTask:
    """train_step / valid_step / evaluation / after_evaluation of lavis/tasks/moment_retrieval.py."""

    def train_step(self, model, samples):
        return model(samples)["loss"]

    def valid_step(self, model, samples, **gen_kwargs):
        out = model.generate(samples, **gen_kwargs)
        assert len(out["qid"]) == len(out["answer"]) == len(out["prediction"])
        return [{"qid": str(q) + "_" + str(i), "raw_prediction": rp, "prediction": p, "target": a, "duration": d}
                for i, (a, q, p, rp, d) in enumerate(zip(out["answer"], out["qid"], out["prediction"], out["raw_prediction"],
                                                         out["duration"]))]

    def evaluation(self, model, loader, device, **gen_kwargs):
        results = []
        for i, samples in enumerate(loader):
            samples = prepare_sample(samples, device)
            samples["iters"] = i
            results.extend(self.valid_step(model, samples, **gen_kwargs))
        return results

    def after_evaluation(self, results, split_name, epoch, result_dir, rank=0, world=1):
        """Per-rank result files concatenated by rank 0 (base_task.py:250-288; like the reference's task, which passes no
        remove_duplicate key, the clips DistributedSampler repeats to even out the ranks are counted twice), then the
        metrics of _report_metrics.  Returns the metrics on rank 0, None elsewhere."""
        os.makedirs(result_dir, exist_ok=True)
        name = "%s_epoch%s" % (split_name, epoch)
        with open(os.path.join(result_dir, "%s_rank%d.json" % (name, rank)), "w") as f:
            json.dump(_plain(results), f)
        mdist.barrier()
        if rank != 0:
            return None
        merged = []
        for r in range(world):
            merged += json.load(open(os.path.join(result_dir, "%s_rank%d.json" % (name, r))))
        path = os.path.join(result_dir, name + ".json")
        with open(path, "w") as f:
            json.dump(merged, f)
        metrics = mr_eval.report_metrics(path)
        logging.info(metrics)
        return metrics


def _plain(x):
    if torch.is_tensor(x):
        return x.tolist()
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


def train_epoch(task, model, loader, optimizer, lr_sched, epoch, device, accum_grad_iters=1, reducer=None, log_freq=50):
    """One epoch of _train_inner_loop (moment_retrieval.py:154-257).  Returns {"loss": mean, "lr": last}."""
    model.train()
    iters = len(loader)
    total, last = 0.0, 0.0
    optimizer.zero_grad()
    for i, samples in enumerate(loader):
        samples = prepare_sample(samples, device)
        samples.update({"epoch": epoch, "num_iters_per_epoch": iters, "iters": i})
        lr_sched.step(cur_epoch=epoch, cur_step=i)
        loss = task.train_step(model, samples)
        loss.backward()
        if (i + 1) % accum_grad_iters == 0:
            if reducer is not None:
                reducer()                                  # one all-reduce(AVG) of the accumulated gradients
            optimizer.step()
            optimizer.zero_grad()
        last = loss.item()
        total += last
        if i % log_freq == 0:
            logging.info("Train: data epoch: [%d]  [%d/%d]  lr: %.6f  loss: %.4f", epoch, i, iters,
                         optimizer.param_groups[0]["lr"], last)
    return {"loss": total / max(iters, 1), "lr": optimizer.param_groups[0]["lr"]}


def build_datasets(cfg, splits, uint8=True, reader=data.Cv2VideoReader):
    """{"train": dataset, "val": dataset, ...} from the recipe's single dataset section (build_info.annotations.<split>.storage,
    build_info.videos.storage, vis_processor.<train|eval>.{n_frms, image_size}) -- base_dataset_builder.py:180-235."""
    (name, ds), = cfg.datasets_cfg.items()
    info = ds["build_info"]
    out = {}
    for split in splits:
        is_train = split == "train"
        vp = ds.get("vis_processor", {}).get("train" if is_train else "eval", {})
        proc = data.VideoProcessor(image_size=vp.get("image_size", 224), n_frms=vp.get("n_frms", 60),
                                   sampling="random" if is_train else "uniform", uint8=uint8, reader=reader, augment=is_train,
                                   min_scale=vp.get("min_scale", 0.5), max_scale=vp.get("max_scale", 1.0))
        out[split] = data.MomentRetrievalDataset(proc, None, info["videos"]["storage"], [info["annotations"][split]["storage"]])
    return out


def build_loader(dataset, batch_size, num_workers, is_train, rank, world, seed=0):
    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=is_train, seed=seed)
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=torch.cuda.is_available(),
                                       sampler=sampler,
                                       shuffle=(sampler is None and is_train), collate_fn=dataset.collater, drop_last=is_train)


def setup_seeds(seed, rank):
    seed = seed + rank                                     # train.py:48-56
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Mr. BLIP stand-alone training")
    ap.add_argument("--cfg-path", required=True)
    ap.add_argument("--options", nargs="+", default=None, help="key=value overrides, e.g. run.batch_size_train=2")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(levelname)s] %(message)s")
    rank, world, local = mdist.init_distributed_mode()
    cfg = Config(args.cfg_path, args.options)
    run = cfg.run_cfg
    setup_seeds(int(run.get("seed", 42)), rank)
    device = torch.device("cuda", local)
    model = registry.get_model_class(cfg.model_cfg["arch"]).from_config(cfg.model_cfg).to(device)
    task = MomentRetrievalTask()
    evaluate_only = bool(run.get("evaluate", False))
    splits = ([] if evaluate_only else list(run.get("train_splits", ["train"]))) + list(run.get("valid_splits", ["val"]))
    datasets = build_datasets(cfg, splits)
    out_dir = run.get("output_dir", "result/mr_BLIP")
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    # MomentRetrievalTask.valid_step calls model.generate(samples) with NO keyword arguments (moment_retrieval.py:33-36), so the
    # reference always decodes with generate's own defaults (num_beams 5, max_length 50, min_length 1) and ignores run.max_len /
    # run.min_len / run.num_beams of the recipe.  Same here; run.generate_from_run_cfg: true opts into forwarding them.
    gen = {}
    if run.get("generate_from_run_cfg", False):
        gen = dict(num_beams=int(run.get("num_beams", 5)), max_length=int(run.get("max_len", 50)), min_length=int(run.get("min_len", 1)))
    loaders = {s: build_loader(d, int(run.get("batch_size_train" if s == "train" else "batch_size_eval", 1)),
                               int(run.get("num_workers", 4)), s == "train", rank, world, int(run.get("seed", 42)))
               for s, d in datasets.items()}
    start_epoch, best, best_epoch = 0, -1.0, -1
    optimizer = lr_sched = reducer = None
    if not evaluate_only:
        optimizer = optim.build_optimizer(model, run["init_lr"], run.get("weight_decay", 0.05), beta2=run.get("beta2", 0.999))
        lr_sched = optim.build_lr_scheduler(optimizer, run)
        if world > 1:
            reducer = mdist.GradAllReducer([p for p in model.parameters() if p.requires_grad], flat_fn=getattr(model, "flat_grads", None))
        if run.get("resume_ckpt_path"):
            start_epoch = checkpoint.resume_checkpoint(model, optimizer, run["resume_ckpt_path"], map_location=device)
    for epoch in range(start_epoch, 1 if evaluate_only else int(run["max_epoch"])):
        if not evaluate_only:
            sampler = getattr(loaders["train"], "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)
            stats = train_epoch(task, model, loaders["train"], optimizer, lr_sched, epoch, device,
                                int(run.get("accum_grad_iters", 1)), reducer)
            if rank == 0:
                with open(os.path.join(out_dir, "log.txt"), "a") as f:
                    f.write(json.dumps({"train_" + k: v for k, v in stats.items()}) + "\n")
        for split in run.get("valid_splits", ["val"]):
            model.eval()
            with torch.no_grad():
                results = task.evaluation(model, loaders[split], device, **gen)
            metrics = task.after_evaluation(results, split, epoch, os.path.join(out_dir, "result"), rank, world)
            if rank == 0 and metrics is not None and not evaluate_only:
                is_best = metrics["agg_metrics"] > best
                if is_best:
                    best, best_epoch = metrics["agg_metrics"], epoch
                    checkpoint.save_checkpoint(model, optimizer, out_dir, epoch, is_best=True, config=cfg)
                with open(os.path.join(out_dir, "log.txt"), "a") as f:
                    f.write(json.dumps({"%s_%s" % (split, k): v for k, v in {**metrics, "best_epoch": best_epoch}.items()}) + "\n")
        if rank == 0 and not evaluate_only:
            checkpoint.save_checkpoint(model, optimizer, out_dir, epoch, config=cfg)
        mdist.barrier()
    mdist.cleanup()


if __name__ == "__main__":
    main()
