"""CPU: the C-ABI library builds, loads and exports every symbol include/sedk.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    return g.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sedk.h")).read()
    return sorted(set(re.findall(r"SEDK_API\s+[\w\s\*]+?\b(sedk_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from desed_task_b200 import _lib
    assert header_symbols() == sorted(_lib.exported_symbols())


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    for name in header_symbols():
        assert hasattr(handle, name), name


def test_plan_struct_layout(built):
    from desed_task_b200 import _lib
    assert _lib.lib().sedk_sizeof_crnn_plan() == ctypes.sizeof(_lib.CrnnPlan)
    assert _lib.lib().sedk_version() >= 100


def test_errors_are_reported_not_swallowed(built):
    from desed_task_b200 import _lib
    L = _lib.lib()
    rc = L.sedk_minmax_init(None, 0, None)
    assert rc == -1
    assert b"sedk_minmax_init" in L.sedk_last_error()
    with pytest.raises(_lib.SedkError):
        _lib.check(rc, "sedk_minmax_init")


def test_cpu_tensors_are_rejected_loudly(built):
    import torch
    from desed_task_b200 import _lib
    from desed_task_b200.frontend import MelSpectrogram
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1)
    with pytest.raises(_lib.SedkError):
        mel(torch.zeros(1, 16000))


def test_front_end_tables_match_oracle():
    import torch
    from oracle import frontend as ofe
    from desed_task_b200.frontend import MelSpectrogram, sparse_filterbank
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1)
    assert torch.equal(mel.mel_scale.fb, ofe.melscale_fbanks())
    assert torch.equal(mel.spectrogram.window, ofe.hamming_window())
    assert set(mel.state_dict().keys()) == {"spectrogram.window", "mel_scale.fb"}
    st, ln, of, w = sparse_filterbank(mel.mel_scale.fb)
    assert int(ln.sum()) == 2024 and int(st.min()) == 1 and int((st + ln).max()) == 1024
    dense = torch.zeros(1025, 128)
    for m in range(128):
        dense[st[m]:st[m] + ln[m], m] = w[of[m]:of[m] + ln[m]]
    assert torch.equal(dense, mel.mel_scale.fb)


def test_unsupported_front_end_configs_raise():
    import torch
    from desed_task_b200.frontend import MelSpectrogram
    with pytest.raises(NotImplementedError):
        MelSpectrogram(16000, 1024, 1024, 256, n_mels=64, power=1)
    with pytest.raises(NotImplementedError):
        MelSpectrogram(16000, 2048, 2048, 256, n_mels=128, power=2.0)


def test_kernel_variant_switches(built, monkeypatch):
    """sedk_set_option / sedk_get_option (include/sedk.h): explicit values win, unset names fall back to the environment
    variable SEDK_<NAME> and then to the caller's default; no CUDA call is involved."""
    from desed_task_b200 import _lib
    L = _lib.lib()
    assert L.sedk_get_option(b"never_set_by_anyone", 7) == 7
    assert L.sedk_get_option(b"never_set_by_anyone", 3) == 7          # the first lookup pins the value
    monkeypatch.setenv("SEDK_FROM_THE_ENVIRONMENT", "5")
    assert L.sedk_get_option(b"from_the_environment", 1) == 5
    assert L.sedk_set_option(b"gru_v2", 0) == 0 and L.sedk_get_option(b"gru_v2", 1) == 0
    assert L.sedk_set_option(b"gru_v2", 1) == 0 and L.sedk_get_option(b"gru_v2", 0) == 1
    assert L.sedk_set_option(None, 1) == -1
