"""Data-parallel plumbing: clips shard across ranks, one gradient all-reduce per step.

The reference has no multi-GPU path at all (train_sed.py:269-276 raises for >1 GPU); this is the one strategy the B200
build adds.  Every rank keeps the per-dataset proportions of the batch ([n_strong, n_weak, n_unlabelled]) so the
index-based masks of training_step stay valid; BatchNorm statistics stay per-rank (no SyncBN in the reference); the
teacher EMA is computed redundantly per rank from identical student weights, so it needs no communication.
The all-reduce is a SUM; the 1/world scale is folded into the fused Adam kernel (grad_scale).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """One process per GPU, launched by torch.distributed.run (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the env)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_batch_sizes(global_batch_sizes, world):
    """[n_strong, n_weak, n_unlabelled] of the GLOBAL batch -> per-rank sizes (must divide evenly: the sub-batch layout
    is positional, sed_trainer.py:286-289)."""
    out = []
    for n in global_batch_sizes:
        if n % world != 0:
            raise ValueError("sub-batch of %d clips does not split over %d ranks" % (n, world))
        out.append(n // world)
    return out


def shard_clip_range(n_clips, rank, world):
    """Contiguous slice of an inference set for this rank (config 5: pure sharding, no collective)."""
    per = (n_clips + world - 1) // world
    lo = min(n_clips, rank * per)
    return lo, min(n_clips, lo + per)


def allreduce_sum_(flat, group=None):
    """In-place SUM of the flat gradient buffer over the group (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_scale(world):
    return 1.0 / float(world)


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)
