"""Stand-in for torch-scatter 2.0.7 (absent in this image) so the REFERENCE modules can be
imported by oracle/pin_against_reference.py.  Test infrastructure only."""
import torch


def scatter_max(src, index, dim=0):
    n = int(index.max()) + 1
    out = torch.full((n,) + tuple(src.shape[1:]), -float("inf"), dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    arg = torch.zeros_like(out, dtype=torch.long)
    return out, arg


def scatter_mean(src, index, dim=0):
    n = int(index.max()) + 1
    s = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
    c = torch.zeros((n,), dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(src[:, 0]))
    return s / c.clamp_min(1).view(-1, *([1] * (src.dim() - 1)))
