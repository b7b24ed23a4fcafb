"""CPU, world_size 2, gloo: host-side logic of the data-parallel path (sharding, gradient averaging convention)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from desed_task_b200 import ddp
    r, l, w = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)
    mine = flat.clone()
    ddp.allreduce_sum_(flat)
    avg = flat * ddp.grad_scale(w)
    # every rank must hold the same averaged gradient
    gathered = [torch.zeros_like(avg) for _ in range(w)]
    dist.all_gather(gathered, avg)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    lin = torch.nn.Linear(4, 4)
    torch.manual_seed(rank)
    torch.nn.init.normal_(lin.weight)
    ddp.broadcast_parameters(lin, 0)
    wts = [torch.zeros_like(lin.weight) for _ in range(w)]
    dist.all_gather(wts, lin.weight.data)
    q.put((rank, same, mine.sum().item(), flat.sum().item(), torch.equal(wts[0], wts[1])))
    dist.destroy_process_group()


def test_allreduce_mean_and_broadcast_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res) and all(r[4] for r in res)
    total = res[0][2] + res[1][2]
    assert abs(res[0][3] - total) < 1e-3 and abs(res[1][3] - total) < 1e-3


def test_sharding_helpers():
    from desed_task_b200 import ddp
    assert ddp.shard_batch_sizes([96, 96, 192], 8) == [12, 12, 24]
    with pytest.raises(ValueError):
        ddp.shard_batch_sizes([12, 12, 25], 2)
    spans = [ddp.shard_clip_range(100000, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 100000
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    assert ddp.shard_clip_range(3, 7, 8) == (3, 3)
    assert ddp.grad_scale(8) == 0.125


def test_scheduler_and_optimizer_host_logic():
    """ExponentialWarmup mirror == oracle formula; FusedAdam bias-correction scalars."""
    from oracle import trainer as otr
    from desed_task_b200.optim import FusedAdam, ema_alpha
    from desed_task_b200.utils.schedulers import ExponentialWarmup
    lin = torch.nn.Linear(3, 3)
    opt = FusedAdam(lin, 1e-3)
    sch = ExponentialWarmup(opt, 1e-3, 1000)
    assert sch.step_num == 1
    for step in (1, 10, 500, 1000, 2000):
        sch.step_num = step
        assert abs(sch._get_scaling_factor() - otr.warmup_scale(step, 1000)) < 1e-12
    sch.step_num = 1
    sch.step()
    assert sch.step_num == 2 and abs(opt.param_groups[0]["lr"] - 1e-3 * otr.warmup_scale(2, 1000)) < 1e-15
    h = opt.hyper(3, 0.5, 0.25)
    lr = opt.param_groups[0]["lr"]
    assert abs(h[0] - lr / (1 - 0.9 ** 3)) < 1e-12 and abs(h[1] - (1 - 0.999 ** 3) ** -0.5) < 1e-9
    assert ema_alpha(0.999, 1) == 0.5 and ema_alpha(0.999, 10 ** 6) == 0.999
    sd = sch.state_dict()
    assert "optimizer" not in sd and sd["step_num"] == 2
