"""TorchScaler on libsedk kernels.  Mirror of desed_task/utils/scaler.py:5-120 (same constructor, `fit`, `forward`,
custom `load_state_dict`, same assertions / NotImplementedError behaviour)."""
import torch

from .._lib import check, lib, ptr, require_cuda, stream_ptr
from ..frontend import new_minmax


class TorchScaler(torch.nn.Module):
    def __init__(self, statistic="dataset", normtype="standard", dims=(1, 2), eps=1e-8):
        super(TorchScaler, self).__init__()
        assert statistic in ["dataset", "instance", None]
        assert normtype in ["standard", "mean", "minmax", None]
        if statistic == "dataset" and normtype == "minmax":
            raise NotImplementedError("statistic==dataset and normtype==minmax is not currently implemented.")
        self.statistic = statistic
        self.normtype = normtype
        self.dims = dims
        self.eps = eps

    def load_state_dict(self, state_dict, strict=True):
        if self.statistic == "dataset":
            super(TorchScaler, self).load_state_dict(state_dict, strict)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        if self.statistic == "dataset":
            super(TorchScaler, self)._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                                           unexpected_keys, error_msgs)

    def fit(self, dataloader, transform_func=lambda x: x[0]):
        """scaler.py:60-88: running mean of per-batch mean and mean-square over `dims` (+ batch)."""
        indx = 0
        for batch in dataloader:
            feats = transform_func(batch)
            m = torch.mean(feats, self.dims, keepdim=True).mean(0).unsqueeze(0)
            m2 = torch.mean(feats ** 2, self.dims, keepdim=True).mean(0).unsqueeze(0)
            if indx == 0:
                mean, mean_squared = m, m2
            else:
                mean += m
                mean_squared += m2
            indx += 1
        mean /= indx
        mean_squared /= indx
        self.register_buffer("mean", mean)
        self.register_buffer("mean_squared", mean_squared)

    def _instance_dims_ok(self, tensor):
        dims = tuple(d % tensor.dim() for d in self.dims)
        return tuple(sorted(dims)) == tuple(range(1, tensor.dim()))

    def forward(self, tensor):
        if self.statistic is None or self.normtype is None:
            return tensor
        require_cuda(tensor)
        x = tensor.float().contiguous()
        B = x.shape[0]
        n = x.numel() // B
        out = torch.empty_like(x)
        s = stream_ptr()
        if self.statistic == "dataset":
            assert hasattr(self, "mean") and hasattr(self, "mean_squared"), \
                "TorchScaler should be fit before used if statistics=dataset"
            assert tensor.ndim == self.mean.ndim, "Pre-computed statistics "
            mean = self.mean.to(x.device).expand(1, *x.shape[1:]).contiguous()
            if self.normtype == "mean":
                check(lib().sedk_affine_bcast(ptr(x), ptr(out), ptr(mean), 0, 1, None, 0, 0, B, n, s),
                      "sedk_affine_bcast")
                return out
            elif self.normtype == "standard":
                std = torch.sqrt(self.mean_squared - self.mean ** 2).to(x.device)
                inv = (1.0 / (std + self.eps)).expand(1, *x.shape[1:]).contiguous()
                check(lib().sedk_affine_bcast(ptr(x), ptr(out), ptr(mean), 0, 1, ptr(inv), 0, 1, B, n, s),
                      "sedk_affine_bcast")
                return out
            raise NotImplementedError
        if not self._instance_dims_ok(x):
            raise NotImplementedError("instance statistics are implemented for dims covering every non-batch axis "
                                      "(the recipes use dims=(1, 2) on [B, n_mels, T]); got dims=%s" % (self.dims,))
        if self.normtype == "minmax":
            mm = new_minmax(B, x.device)
            # min / max pass (identity copy fused with the reduction), then the affine map
            check(lib().sedk_feat_mix_log(ptr(x), None, None, ptr(out), B, n, 0, 0.0, 0.0, 0.0, ptr(mm), s),
                  "sedk_feat_mix_log")
            check(lib().sedk_minmax_scale(ptr(x), ptr(out), ptr(mm), B, n, self.eps, s), "sedk_minmax_scale")
            return out
        stats = torch.empty(B, 2, device=x.device, dtype=torch.float32)
        check(lib().sedk_instance_stats(ptr(x), ptr(stats), B, n, s), "sedk_instance_stats")
        mean = stats[:, 0].contiguous()
        if self.normtype == "mean":
            check(lib().sedk_affine_bcast(ptr(x), ptr(out), ptr(mean), 1, 0, None, 0, 0, B, n, s), "sedk_affine_bcast")
            return out
        inv = (1.0 / (stats[:, 1] + self.eps)).contiguous()
        check(lib().sedk_affine_bcast(ptr(x), ptr(out), ptr(mean), 1, 0, ptr(inv), 1, 0, B, n, s), "sedk_affine_bcast")
        return out
