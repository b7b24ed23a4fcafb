"""The fused NVLink all-reduce + EMA/Adam kernel (csrc/nvls.cu) on ONE GPU: world = 1 runs the same code path (flag
hand-shake with itself, tile ownership, P2P branch) and must reproduce sedk_adam_ema_dev bit for bit.  The 2-rank / multicast
check is tools/nvls_check.py (torchrun, 2 GPUs; profiles/r2_nvls_check.txt); world-size-2 host logic: tests/test_ddp_cpu.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _call(L, p, m, v, ema, g, flags, hyper, do_adam, stream_ptr, ptr):
    VP = ctypes.c_void_p * 1
    return L.sedk_allreduce_adam_nvls(ptr(p), ptr(m), ptr(v), ptr(ema), g.numel(), do_adam, 0.9, 0.999, 1e-8, ptr(hyper),
                                      None, VP(g.data_ptr()), VP(flags.data_ptr()), 0, 1, stream_ptr())


@pytest.mark.parametrize("n", [4, 2048, 1112420, 1785096])
def test_single_rank_equals_the_plain_fused_update(dev, n):
    from desed_task_b200._lib import check, lib, ptr, stream_ptr
    L = lib()
    gen = torch.Generator(device=dev).manual_seed(n)
    p = torch.randn(n, device=dev, generator=gen)
    g = torch.randn(n, device=dev, generator=gen) * 1e-2
    m = torch.randn(n, device=dev, generator=gen) * 1e-3
    v = torch.rand(n, device=dev, generator=gen) * 1e-4
    ema = torch.randn(n, device=dev, generator=gen)
    hyper = torch.tensor([1e-3 / (1 - 0.9 ** 3), 1.0 / (1 - 0.999 ** 3) ** 0.5, 0.999, 0.5], device=dev)
    ref = [t.clone() for t in (p, m, v, ema)]
    check(L.sedk_adam_ema_dev(ptr(ref[0]), ptr(g), ptr(ref[1]), ptr(ref[2]), ptr(ref[3]), n, 1, 0.9, 0.999, 1e-8, ptr(hyper),
                              stream_ptr()), "sedk_adam_ema_dev")
    flags = torch.zeros(int(L.sedk_nvls_flag_bytes()) // 4, dtype=torch.int32, device=dev)
    g0 = g.clone()
    for rep in range(3):                      # launch epochs: repeated launches must keep working
        got = [t.clone() for t in (p, m, v, ema)]
        check(_call(L, got[0], got[1], got[2], got[3], g, flags, hyper, 1, stream_ptr, ptr), "sedk_allreduce_adam_nvls")
        torch.cuda.synchronize()
        assert torch.equal(g, g0)             # the sum over one rank
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    # all-reduce only: nothing but the gradient buffer is touched
    got = [t.clone() for t in (p, m, v, ema)]
    check(_call(L, got[0], got[1], got[2], got[3], g, flags, hyper, 0, stream_ptr, ptr), "sedk_allreduce_adam_nvls")
    torch.cuda.synchronize()
    for a, b in zip(got, (p, m, v, ema)):
        assert torch.equal(a, b)


def test_bad_arguments_are_refused(dev):
    from desed_task_b200._lib import lib, ptr, stream_ptr
    L = lib()
    t = torch.zeros(8, device=dev)
    flags = torch.zeros(int(L.sedk_nvls_flag_bytes()) // 4, dtype=torch.int32, device=dev)
    VP = ctypes.c_void_p * 1
    args = (ptr(t), ptr(t), ptr(t), None)
    assert L.sedk_allreduce_adam_nvls(*args, 6, 1, 0.9, 0.999, 1e-8, ptr(t), None, VP(t.data_ptr()), VP(flags.data_ptr()),
                                      0, 1, stream_ptr()) != 0         # n not a multiple of 4
    assert L.sedk_allreduce_adam_nvls(*args, 8, 1, 0.9, 0.999, 1e-8, ptr(t), None, VP(t.data_ptr()), VP(flags.data_ptr()),
                                      0, 9, stream_ptr()) != 0         # more than one NVSwitch domain
    assert L.sedk_allreduce_adam_nvls(*args, 8, 1, 0.9, 0.999, 1e-8, None, None, VP(t.data_ptr()), VP(flags.data_ptr()),
                                      0, 1, stream_ptr()) != 0         # fused update without the device scalars
