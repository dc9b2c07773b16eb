"""Shared plumbing of the drop-in modules: parameter containers keep the
reference's names and layouts (so its checkpoints load unchanged), while the
forward pass feeds GEMM-ready copies of the weights to the CUDA kernels."""
import torch
from torch import nn


class WeightCache:
    """Caches `weight.reshape(Cout, -1).T.contiguous()` ([K, Cout], the layout
    rf_conv3d_fwd / rf_linear_fwd read) and refreshes it when the parameter is
    modified in place (optimizer step, load_state_dict) or moved."""

    def __init__(self):
        self._store = {}

    def wt(self, p: torch.Tensor) -> torch.Tensor:
        key = id(p)
        tag = (p.data_ptr(), p._version, p.device)
        hit = self._store.get(key)
        if hit is None or hit[0] != tag:
            t = p.detach().reshape(p.shape[0], -1).t().contiguous()
            self._store[key] = (tag, t)
            return t
        return hit[1]

    def derived(self, name, params, fn):
        """Caches fn(*params) under `name`, invalidated when any param changes."""
        tag = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self._store.get(name)
        if hit is None or hit[0] != tag:
            with torch.no_grad():
                t = fn(*[p.detach() for p in params])
            if hit is not None:  # the old image is freed: captured CUDA graphs that read it are stale
                from .. import ops
                ops.bump_generation()
            self._store[name] = (tag, t)
            return t
        return hit[1]


class RfModule(nn.Module):
    """nn.Module whose forward runs on the rf_b200 kernels (CUDA only)."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_wcache", WeightCache())

    def _wt(self, p):
        return self._wcache.wt(p)
