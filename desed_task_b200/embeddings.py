"""Embedding storage format of the 2024 recipe on the device.

Upstream stores BEATs frame embeddings as fp32 [768, 496] per clip (1.52 MB: recipes/dcase2024_task4_baseline/
extract_embeddings.py:48-53, desed_task/dataio/datasets.py:221-228) and `CRNN.forward` pools them to its 156 frames on every
call (desed_task/nnet/CRNN.py:280-283).  `pool_embeddings` does that aggregation ONCE with the arithmetic of the fusion
kernel and stores [B, 768, 156] as fp32 or bf16 (240 KB per clip, 6.3x less H2D / storage); a pre-pooled tensor is fed to
`CRNN(..., embeddings=...)` / the engines like any other embedding tensor (pooling 156 -> 156 frames is the identity, in the
reference too).  fp32 pre-pooling is exact; bf16 storage rounds the embeddings to 8 significant bits (opt-in)."""
import torch

from ._lib import check, lib, ptr, require_cuda, stream_ptr


def pool_embeddings(emb, frames=156, mode="pool1d", dtype=torch.bfloat16):
    """emb: cuda fp32 [B, E, Te] -> [B, E, frames] in `dtype` (torch.float32 or torch.bfloat16)."""
    require_cuda(emb)
    if mode not in ("pool1d", "interpolate") or dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("pool_embeddings: mode in {pool1d, interpolate}, dtype in {float32, bfloat16}")
    x = emb.float().contiguous()
    B, E, Te = x.shape
    out = torch.empty(B, E, frames, device=x.device, dtype=dtype)
    check(lib().sedk_pool_embeddings(ptr(x), ptr(out), B, E, Te, frames, 1 if mode == "interpolate" else 0,
                                     1 if dtype == torch.bfloat16 else 0, stream_ptr()), "sedk_pool_embeddings")
    return out


def upcast(emb_bf16, out=None):
    """bf16 embeddings -> the fp32 working precision of the fusion kernels."""
    require_cuda(emb_bf16)
    x = emb_bf16.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    check(lib().sedk_bf16_to_f32(ptr(x), ptr(out), x.numel(), stream_ptr()), "sedk_bf16_to_f32")
    return out
