"""ctypes binding of libsedk.so (the C-ABI boundary, include/sedk.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsedk.so")

SEDK_MAX_CONV = 8
SEDK_MAX_GRU_LAYERS = 4

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f32 = C.c_float
u64 = C.c_uint64


class SedkError(RuntimeError):
    pass


class MelTables(C.Structure):
    _fields_ = [("window", vp), ("tw2048", vp), ("tw32x32", vp), ("fb_start", vp), ("fb_len", vp), ("fb_off", vp),
                ("fb_w", vp), ("n_mels", C.c_int32), ("hop", C.c_int32)]


class ConvLayer(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("T", C.c_int32), ("F", C.c_int32), ("pt", C.c_int32),
                ("pf", C.c_int32),
                ("w", vp), ("b", vp), ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("num_batches", vp), ("glu_w", vp), ("glu_b", vp),
                ("gw", vp), ("gb", vp), ("ggamma", vp), ("gbeta", vp), ("gglu_w", vp), ("gglu_b", vp),
                ("wpack", vp), ("gwpack", vp), ("z", vp), ("gy", vp), ("out", vp), ("gout", vp), ("stats", vp),
                ("bn", vp), ("glu_pack", vp), ("lin", vp)]


class GruLayer(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("hidden", C.c_int32),
                ("w_ih", vp * 2), ("w_hh", vp * 2), ("b_ih", vp * 2), ("b_hh", vp * 2),
                ("gw_ih", vp * 2), ("gw_hh", vp * 2), ("gb_ih", vp * 2), ("gb_hh", vp * 2),
                ("gi", vp * 2), ("gates", vp * 2), ("hprev", vp * 2), ("dghn", vp * 2),
                ("out", vp), ("gout", vp)]


class CrnnPlan(C.Structure):
    _fields_ = [("B", C.c_int32), ("n_mels", C.c_int32), ("n_frames", C.c_int32), ("n_conv", C.c_int32),
                ("n_gru", C.c_int32), ("nclass", C.c_int32), ("training", C.c_int32), ("precision", C.c_int32),
                ("bn_eval", C.c_int32), ("activation", C.c_int32),
                ("dropout_p", f32), ("bn_eps", f32), ("bn_momentum", f32), ("seed", u64), ("seed_dev", vp),
                ("x", vp), ("x_sb", i64), ("x_sm", i64), ("x_st", i64), ("minmax", vp), ("scaler_eps", f32),
                ("specaug", vp), ("x0", vp),
                ("conv", ConvLayer * SEDK_MAX_CONV), ("gru", GruLayer * SEDK_MAX_GRU_LAYERS),
                ("emb", vp), ("emb_dim", C.c_int32), ("emb_T", C.c_int32), ("emb_mode", C.c_int32), ("cat_w", vp), ("cat_b", vp),
                ("gcat_w", vp), ("gcat_b", vp), ("cat_in", vp), ("fused", vp), ("gfused", vp), ("dropstep", vp),
                ("dense_w", vp), ("dense_b", vp), ("soft_w", vp), ("soft_b", vp),
                ("gdense_w", vp), ("gdense_b", vp), ("gsoft_w", vp), ("gsoft_b", vp),
                ("classes_mask", vp), ("rnn_drop", vp), ("grnn_drop", vp), ("strong", vp), ("weak", vp), ("sof", vp), ("hsum", vp),
                ("gstrong", vp), ("gweak", vp),
                ("zero_fwd", vp), ("zero_fwd_bytes", i64), ("zero_bwd", vp), ("zero_bwd_bytes", i64), ("l0_sums", vp)]


_SIGS = {
    "sedk_version": (i32, []),
    "sedk_device_cc": (i32, []),
    "sedk_sizeof_crnn_plan": (i32, []),
    "sedk_launch_count": (C.c_longlong, []),
    "sedk_set_tcgen05": (i32, [i32]),
    "sedk_set_gru_cluster": (i32, [i32]),
    "sedk_get_tcgen05": (i32, []),
    "sedk_set_option": (i32, [C.c_char_p, i32]),
    "sedk_get_option": (i32, [C.c_char_p, i32]),
    "sedk_profile_enable": (i32, [i32]),
    "sedk_profile_report": (i32, [C.c_char_p, i32]),
    "sedk_logmel_fwd": (i32, [vp, i32, i32, C.POINTER(MelTables), vp, i64, i64, i64, i32, f32, f32, f32, vp, vp]),
    "sedk_logmel_fwd_i16": (i32, [vp, i32, i32, C.POINTER(MelTables), vp, i64, i64, i64, i32, f32, f32, f32, vp, vp]),
    "sedk_minmax_init": (i32, [vp, i32, vp]),
    "sedk_minmax_decode": (i32, [vp, vp, i32, vp]),
    "sedk_feat_mix_log": (i32, [vp, vp, vp, vp, i32, i64, i32, f32, f32, f32, vp, vp]),
    "sedk_minmax_scale": (i32, [vp, vp, vp, i32, i64, f32, vp]),
    "sedk_instance_stats": (i32, [vp, vp, i32, i64, vp]),
    "sedk_affine_bcast": (i32, [vp, vp, vp, i64, i64, vp, i64, i64, i32, i64, vp]),
    "sedk_label_mix": (i32, [vp, vp, vp, vp, i32, i64, i32, vp]),
    "sedk_roll_last": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "sedk_add_noise": (i32, [vp, vp, vp, vp, vp, i32, i64, vp]),
    "sedk_adam_ema": (i32, [vp, vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, i32, f32, f32, vp]),
    "sedk_adam_ema_dev": (i32, [vp, vp, vp, vp, vp, i64, i32, f32, f32, f32, vp, vp]),
    "sedk_nvls_flag_bytes": (i64, []),
    "sedk_nvls_debug_stamps": (i32, [vp]),
    "sedk_nvls_fault": (i32, []),
    "sedk_allreduce_adam_nvls": (i32, [vp, vp, vp, vp, i64, i32, f32, f32, f32, vp, vp, vp, vp, i32, i32, vp]),
    "sedk_bump_counter": (i32, [vp, u64, vp]),
    "sedk_sumsq": (i32, [vp, i64, vp, vp]),
    "sedk_mask_spans": (i32, [vp, i32, i32, i32, i32, i32, u64, vp, u64, vp]),
    "sedk_median_filter": (i32, [vp, vp, i32, i32, i32, i64, i64, i64, i64, i64, i64, vp, vp]),
    "sedk_pool_embeddings": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "sedk_bf16_to_f32": (i32, [vp, vp, i64, vp]),
    "sedk_encode_strong": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "sedk_decode_events": (i32, [vp, i32, i32, i32, i64, i64, i64, vp, i32, vp, vp, vp, i32, vp]),
    "sedk_crnn_forward": (i32, [C.POINTER(CrnnPlan), vp]),
    "sedk_crnn_backward": (i32, [C.POINTER(CrnnPlan), vp]),
    "sedk_crnn_backward_phase": (i32, [C.POINTER(CrnnPlan), i32, vp]),
    "sedk_sed_loss": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp]),
    "sedk_sed_loss_dev": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "sedk_sed_loss_ex": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp]),
    "sedk_conv3x3": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "sedk_conv_wgrad": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "sedk_gemm": (i32, [i32, i32, i32, i32, i32, f32, vp, i32, vp, i32, f32, vp, i32, vp, i32, vp]),
}

_lib = None


def exported_symbols():
    """Every symbol include/sedk.h declares (used by the CPU-side load test)."""
    return ["sedk_last_error"] + list(_SIGS)


def lib():
    """Load libsedk.so (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SedkError(
            "libsedk.so is not built (%s). Build it with `python -m desed_task_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback." % LIB_PATH)
    handle = C.CDLL(LIB_PATH)
    handle.sedk_last_error.restype = C.c_char_p
    handle.sedk_last_error.argtypes = []
    for name, (res, args) in _SIGS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    if handle.sedk_sizeof_crnn_plan() != C.sizeof(CrnnPlan):
        raise SedkError("sedk_crnn_plan layout mismatch: C %d bytes vs ctypes %d bytes"
                        % (handle.sedk_sizeof_crnn_plan(), C.sizeof(CrnnPlan)))
    _lib = handle
    return handle


def check(rc, what=""):
    if rc != 0:
        msg = lib().sedk_last_error().decode("utf-8", "replace")
        raise SedkError("%s failed (rc=%d): %s" % (what or "libsedk call", rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SedkError("desed_task_b200 ops run on CUDA tensors only (got a %s tensor); "
                            "there is no CPU fallback" % t.device)
