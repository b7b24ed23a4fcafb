"""Symmetric-memory plumbing for the fused NVLink all-reduce + Adam kernel (csrc/nvls.cu, include/sedk.h
sedk_allreduce_adam_nvls): the flat gradient buffer and a flag block are allocated with torch.distributed._symmetric_memory
(cuMem allocations mapped into every peer, one multicast address when the fabric has NVLS), exchanged once at start-up.
PyTorch is plumbing here (allocation + handle exchange); the collective itself is this library's kernel - no NCCL call on
the step.  The reference has no multi-GPU path (train_sed.py:269-276)."""
import ctypes
import sys

import torch
import torch.distributed as dist

from ._lib import check, lib, ptr, stream_ptr


class NvlsGradient:
    """Owns the symmetric gradient region of one rank.  `alloc(n)` is handed to the CRNN workspace as its gradient
    allocator (the first n_grad floats of that region are the flat gradient, parameters() order); `connect()` is the
    collective handle exchange; `step()` launches the fused kernel on the current stream."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise RuntimeError("the fused NVLink all-reduce covers one NVSwitch domain (<= 8 ranks)")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = None
        self.flags = symm.empty(int(lib().sedk_nvls_flag_bytes()) // 4, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self.connected = False
        self.multicast = False

    def alloc(self, numel):
        if self.buf is not None:
            if self.buf.numel() == numel:
                return self.buf
            # another batch shape of the same module (e.g. validation): ordinary memory, that workspace is not reduced here
            return torch.zeros(int(numel), dtype=torch.float32, device=self.device)
        self.buf = self.symm.empty(int(numel), dtype=torch.float32, device=self.device)
        self.buf.zero_()
        return self.buf

    def connect(self, use_multicast=True):
        """Collective: every rank calls it once, after its workspace exists."""
        gname = self.group.group_name if hasattr(self.group, "group_name") else self.group
        hg = self.symm.rendezvous(self.buf, gname)
        hf = self.symm.rendezvous(self.flags, gname)

        def peers(h, t):
            off = t.data_ptr() - int(h.buffer_ptrs[self.rank])
            if off < 0 or off > int(h.buffer_size):
                raise RuntimeError("symmetric-memory handle does not cover the tensor")
            return [int(b) + off for b in h.buffer_ptrs], off

        self.g_ptrs, goff = peers(hg, self.buf)
        self.f_ptrs, _ = peers(hf, self.flags)
        mc = int(getattr(hg, "multicast_ptr", 0) or 0)
        self.g_mc = mc + goff if (mc and use_multicast) else 0
        self.multicast = bool(self.g_mc)
        VP = ctypes.c_void_p * self.world
        self._g_arr = VP(*self.g_ptrs)
        self._f_arr = VP(*self.f_ptrs)
        self._handles = (hg, hf)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)           # every rank's flag block is zero and mapped before the first launch
        self.connected = True
        return self

    def step(self, opt, n, ema_flat, hyper_dev, do_adam=True):
        g = opt.param_groups[0]
        opt._ensure()
        check(lib().sedk_allreduce_adam_nvls(ptr(opt.flat), ptr(opt.m), ptr(opt.v), ptr(ema_flat), int(n), 1 if do_adam else 0,
                                             g["betas"][0], g["betas"][1], g["eps"], ptr(hyper_dev),
                                             ctypes.c_void_p(self.g_mc) if self.g_mc else None, self._g_arr, self._f_arr,
                                             self.rank, self.world, stream_ptr()), "sedk_allreduce_adam_nvls")


def try_create(group=None, device=None):
    """NvlsGradient, or None (with the reason on stderr) where symmetric memory is unavailable.  Collective-safe: the
    decision is agreed over the group, so either every rank gets an object or none does."""
    obj, why = None, ""
    try:
        obj = NvlsGradient(group, device)
    except Exception as e:      # noqa: BLE001 - any failure means "use the NCCL path"
        why = "%s: %s" % (type(e).__name__, e)
    ok = torch.tensor([1 if obj is not None else 0], device=device if device is not None else "cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        if why:
            print("[desed_task_b200] fused NVLink all-reduce unavailable (%s); using NCCL" % why, file=sys.stderr)
        return None
    return obj
