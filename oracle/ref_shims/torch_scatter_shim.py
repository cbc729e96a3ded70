"""torch_scatter.scatter stand-in (test infrastructure; see oracle/ref_shims/__init__.py).

Call site in the reference: creste/utils/projection.py:124 (`reduce='max'`).  torch_scatter
semantics: cells that receive no source are 0; otherwise the reduction of the sources only.
That is exactly `zeros.scatter_reduce(..., include_self=False)`.
"""
import torch

_RED = {"max": "amax", "min": "amin", "sum": "sum", "add": "sum", "mean": "mean"}


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    if index.dim() != src.dim():
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape).expand_as(src)
    size = list(src.shape)
    size[dim] = int(dim_size) if dim_size is not None else int(index.max()) + 1
    base = torch.zeros(size, dtype=src.dtype, device=src.device)
    return base.scatter_reduce(dim, index, src, _RED[reduce], include_self=False)
