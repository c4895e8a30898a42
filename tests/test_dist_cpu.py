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
