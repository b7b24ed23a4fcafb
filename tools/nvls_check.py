"""2..8-rank check + timing of the fused NVLink all-reduce + Adam kernel against NCCL all-reduce + sedk_adam_ema_dev.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/nvls_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desed_task_b200 import ddp, nvls  # noqa: E402
from desed_task_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from desed_task_b200.optim import FusedAdam  # noqa: E402


class Flat(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n))


def main():
    rank, local, world = ddp.init_from_env()
    dev = torch.device("cuda", local)
    n = 1112420
    L = lib()
    out = []
    for use_mc in (True, False):
        obj = nvls.try_create(None, dev)
        if obj is None:
            if rank == 0:
                print("symmetric memory unavailable")
            return
        g = obj.alloc(n + 4096)          # like the workspace: gradient + packed accumulators behind it
        obj.connect(use_multicast=use_mc)
        gen = torch.Generator(device=dev).manual_seed(1)          # same parameters on every rank
        net = Flat(n).to(dev)
        net.w.data.copy_(torch.randn(n, device=dev, generator=gen))
        opt = FusedAdam(net, 1e-3)
        opt._ensure()
        ema = torch.randn(n, device=dev, generator=gen)
        opt.m.copy_(torch.randn(n, device=dev, generator=gen) * 1e-3)
        opt.v.copy_(torch.rand(n, device=dev, generator=gen) * 1e-4)
        hyper = torch.tensor(opt.hyper(3, 0.999, 1.0 / world), device=dev)
        gr = torch.Generator(device=dev).manual_seed(100 + rank)  # different gradients per rank
        grad = torch.randn(n, device=dev, generator=gr) * 1e-2
        # reference: NCCL all-reduce + the plain fused update on clones
        ref_g = grad.clone()
        dist.all_reduce(ref_g)
        ref = [t.clone() for t in (opt.flat, opt.m, opt.v, ema)]
        check(L.sedk_adam_ema_dev(ptr(ref[0]), ptr(ref_g), ptr(ref[1]), ptr(ref[2]), ptr(ref[3]), n, 1, 0.9, 0.999, 1e-8,
                                  ptr(hyper), stream_ptr()), "sedk_adam_ema_dev")
        state0 = [t.clone() for t in (opt.flat, opt.m, opt.v, ema)]
        worst = 0.0
        for rep in range(3):
            for t, s in zip((opt.flat, opt.m, opt.v, ema), state0):
                t.copy_(s)
            g[:n].copy_(grad)
            torch.cuda.synchronize()
            dist.barrier()
            obj.step(opt, n, ema, hyper, do_adam=True)
            torch.cuda.synchronize()
            dg = (g[:n] - ref_g).abs().max().item()
            dp = max((a - b).abs().max().item() for a, b in zip((opt.flat, opt.m, opt.v, ema), ref))
            worst = max(worst, dg, dp)
        # replicas must be bit-identical: compare rank 0's parameters with everyone's
        p0 = opt.flat.clone()
        dist.broadcast(p0, 0)
        same = bool(torch.equal(p0, opt.flat))
        # timing: fused kernel vs NCCL all-reduce + update (device events, 100 launches each)
        def timed(fn, iters=100):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item() * 1e3

        t_fused = timed(lambda: obj.step(opt, n, ema, hyper, do_adam=True))
        t_grid = {}
        for gsz in (32, 64, 256):
            L.sedk_set_option(b"nvls_grid", gsz)
            t_grid[gsz] = round(timed(lambda: obj.step(opt, n, ema, hyper, do_adam=True)), 1)
        L.sedk_set_option(b"nvls_grid", 128)
        L.sedk_set_option(b"nvls_debug", 1)
        import ctypes
        stamps = (ctypes.c_uint64 * 5)()
        phases = []
        for _ in range(5):
            torch.cuda.synchronize()
            dist.barrier()
            obj.step(opt, n, ema, hyper, do_adam=True)
            torch.cuda.synchronize()
            check(L.sedk_nvls_debug_stamps(stamps), "stamps")
            phases.append([round((stamps[i + 1] - stamps[i]) / 1e3, 1) for i in range(4)])
        L.sedk_set_option(b"nvls_debug", 0)

        def nccl_path():
            dist.all_reduce(g[:n])
            check(L.sedk_adam_ema_dev(ptr(opt.flat), ptr(g), ptr(opt.m), ptr(opt.v), ptr(ema), n, 1, 0.9, 0.999, 1e-8,
                                      ptr(hyper), stream_ptr()), "sedk_adam_ema_dev")
        t_nccl = timed(nccl_path)
        t_upd = timed(lambda: check(L.sedk_adam_ema_dev(ptr(opt.flat), ptr(g), ptr(opt.m), ptr(opt.v), ptr(ema), n, 1, 0.9,
                                                        0.999, 1e-8, ptr(hyper), stream_ptr()), "adam"))
        out.append("world %d  multicast %s (requested %s): max |fused - (NCCL + update)| = %.3g, replicas identical: %s | "
                   "fused kernel %.1f us, NCCL all-reduce + update %.1f us, update alone %.1f us | grid 32/64/256: %s | CTA 0 phases "
                   "[barrier A, phase 1, barrier B, phase 2] us: %s" % (
                       world, obj.multicast, use_mc, worst, same, t_fused, t_nccl, t_upd, t_grid, phases[-3:]))
        del obj
    if rank == 0:
        print("\n".join(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
