"""N > 1 host logic on CPU: two gloo ranks average their gradients through GradAllReducer (the only exchange
step of the data-parallel path) and shard clips by rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from mr_blip_b200 import dist as mdist
    r, w, _ = mdist.init_distributed_mode(backend="gloo")
    assert (r, w) == (rank, world) and mdist.is_dist()
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(8, 16)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(3), requires_grad=False)]
    params[0].grad = torch.full((8, 16), float(rank + 1))
    params[1].grad = None                                  # a rank without this gradient contributes zeros
    if rank == 1:
        params[1].grad = torch.arange(5.0)
    red = mdist.GradAllReducer(params)
    red()
    ok = torch.allclose(params[0].grad, torch.full((8, 16), 1.5)) and torch.allclose(params[1].grad, torch.arange(5.0) / 2)
    ok = ok and params[2].grad is None
    # zero-copy path: every .grad is a view of one flat buffer (what BLIP2_MR.flat_grads() hands to the reducer)
    flat = torch.arange(8 * 16 + 5, dtype=torch.float32) * (rank + 1)
    params[0].grad, params[1].grad = flat[:128].view(8, 16), flat[128:]
    mdist.GradAllReducer(params, flat_fn=lambda: flat)()
    want = torch.arange(8 * 16 + 5, dtype=torch.float32) * 1.5
    ok = ok and torch.allclose(flat, want) and torch.allclose(params[0].grad.reshape(-1), want[:128])
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_two_gloo_ranks():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


class _Standin(torch.nn.Module):
    """Stand-in for the model in the loop test: loss pulls w towards the mean duration of the rank's clips."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, samples):
        return {"loss": ((self.w - samples["duration"].float().mean()) ** 2).sum()}

    def generate(self, samples, **kw):
        n = len(samples["query_prompt"])
        return {"prediction": ["[[1, 3]]"] * n, "raw_prediction": ["[[1, 3]]"] * n, "answer": list(samples["relevant_windows"]),
                "qid": samples["query_id"].tolist(), "duration": samples["duration"].tolist()}


class _Clips(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {"video": torch.zeros(2, 3, 4, 4, dtype=torch.uint8), "duration": torch.tensor(float(i)), "query_id": i,
                "timestamps": torch.tensor([0.5, 1.5]), "video_prompt_end": "<extra_id_0>", "query_prompt": "Query: q%d\n" % i,
                "task_prompt": "t", "relevant_windows": "[[1, 3]]"}

    collater = staticmethod(torch.utils.data.dataloader.default_collate)


def _loop_worker(rank, world, port, out, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from mr_blip_b200 import dist as mdist, optim, train
    mdist.init_distributed_mode(backend="gloo")
    ds = _Clips(7)                                                     # odd: the sampler pads one rank with a repeated clip
    loader = train.build_loader(ds, 2, 0, True, rank, world, seed=3)
    val = train.build_loader(ds, 2, 0, False, rank, world)
    model = _Standin()
    opt = torch.optim.SGD(model.parameters(), lr=0.25)
    sched = optim.LinearWarmupCosineLRScheduler(opt, max_epoch=1, min_lr=0.25, init_lr=0.25)
    red = mdist.GradAllReducer(list(model.parameters()))
    task = train.MomentRetrievalTask()
    loader.sampler.set_epoch(0)
    seen = [int(i) for b in loader for i in b["query_id"]]
    loader.sampler.set_epoch(0)
    train.train_epoch(task, model, loader, opt, sched, 0, "cpu", accum_grad_iters=1, reducer=red)
    results = task.evaluation(model.eval(), val, "cpu")
    metrics = task.after_evaluation(results, "val", 0, tmp, rank, world)
    out[rank] = {"w": float(model.w.detach()), "seen": seen, "n_results": len(results),
                 "metrics": None if metrics is None else {"total": metrics["total"], "agg": metrics["agg_metrics"]}}
    dist.barrier()
    dist.destroy_process_group()


def test_training_loop_two_gloo_ranks(tmp_path):
    """train.train_epoch / evaluation / after_evaluation on two gloo ranks: clips sharded by DistributedSampler (disjoint up
    to its padding), one gradient all-reduce per optimiser step keeps the replicas identical, per-rank result files are
    concatenated by rank 0 (padded duplicate included, as in the reference's save_result)."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_loop_worker, args=(2, port, out, str(tmp_path)), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["w"] == b["w"] and a["w"] != 0.0                       # same averaged updates on both replicas
    assert len(a["seen"]) == len(b["seen"]) == 4 and set(a["seen"]) | set(b["seen"]) == set(range(7))   # 7 clips padded to 4 + 4
    assert a["n_results"] + b["n_results"] == 8                      # 7 clips padded to 8 over two ranks
    assert b["metrics"] is None and a["metrics"]["total"] == 8 and a["metrics"]["agg"] == 100.0
