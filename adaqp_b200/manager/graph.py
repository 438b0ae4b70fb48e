"""Device-resident partition graph for the aggregation kernels.

Stands in for the DGL graph objects the reference keeps on the GPU
(AdaQP/manager/graphEngine.py:62-63,149-161): dst-major CSR of all in-edges of the inner
nodes, global degrees in `ndata`, and the degree norms of AdaQP/model/ops.py:21-25,49-57
precomputed once with the same torch expressions (deg.float().clamp(min=1).pow(p)).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib


class LocalGraph(object):
    def __init__(self, indptr: np.ndarray, indices: np.ndarray, in_degrees: np.ndarray,
                 out_degrees: np.ndarray, n_inner: int, n_halo: int, device: torch.device):
        self.device = torch.device(device)
        self.n_inner, self.n_halo = int(n_inner), int(n_halo)
        self.indptr = torch.from_numpy(np.ascontiguousarray(indptr, np.int64)).to(self.device)
        self.indices = torch.from_numpy(np.ascontiguousarray(indices, np.int32)).to(self.device)
        self.nnz = int(self.indices.numel())
        # first halo column of every row (columns are sorted, halo ids >= n_inner): lets the
        # aggregation of a row be split into its local-source and halo-source segments
        ip = np.ascontiguousarray(indptr, np.int64)
        is_local = (np.asarray(indices) < self.n_inner)
        csum = np.concatenate([[0], np.cumsum(is_local, dtype=np.int64)])
        self.halo_split = torch.from_numpy(ip[:-1] + (csum[ip[1:]] - csum[ip[:-1]])).to(self.device)
        self.ndata: Dict[str, torch.Tensor] = {
            "in_degrees": torch.from_numpy(np.ascontiguousarray(in_degrees)).to(self.device),
            "out_degrees": torch.from_numpy(np.ascontiguousarray(out_degrees)).to(self.device),
        }
        ind = self.ndata["in_degrees"].float().clamp(min=1)
        outd = self.ndata["out_degrees"].float().clamp(min=1)
        self.norm = {
            "in_-0.5": ind.pow(-0.5).contiguous(), "out_-0.5": outd.pow(-0.5).contiguous(),
            "out_-1": torch.pow(outd, -1).contiguous(), "in_+1_-1": torch.pow(ind + 1, -1).contiguous(),
            "out_+1_-1": torch.pow(outd + 1, -1).contiguous(),
        }

    def num_nodes(self) -> int:
        return self.n_inner + self.n_halo

    def num_edges(self) -> int:
        return self.nnz

    def to(self, device):
        return self if torch.device(device) == self.device else NotImplemented


def spmm(graph: LocalGraph, x_local: torch.Tensor, x_halo: Optional[torch.Tensor],
         pre: Optional[torch.Tensor], post: Optional[torch.Tensor], mean: bool = False,
         add_self: bool = False, row_begin: int = 0, row_end: Optional[int] = None,
         out: Optional[torch.Tensor] = None, stream=None, part: Optional[str] = None) -> torch.Tensor:
    """out[v - row_begin] = post[v] * sum_u pre[u] x[u]  over the CSR rows [row_begin, row_end).
    part='local': only the local-source neighbours of each row (no halo needed);
    part='halo' : only the halo-source neighbours, ACCUMULATED into `out`."""
    L = _lib.load()
    row_end = graph.n_inner if row_end is None else int(row_end)
    F = int(x_local.shape[1])
    assert x_local.dtype == torch.float32 and x_local.stride(1) == 1
    if out is None:
        out = torch.empty((row_end - row_begin, F), dtype=torch.float32, device=x_local.device)
    if x_halo is not None and x_halo.shape[0] == 0:
        x_halo = None
    seg_start = seg_end = None
    accumulate = 0
    if part == "local":
        seg_end = graph.halo_split.data_ptr()
    elif part == "halo":
        assert out is not None, "the halo part accumulates into the output of the local part"
        seg_start = graph.halo_split.data_ptr()
        accumulate, add_self = 1, False
    elif part is not None:
        raise ValueError(part)
    rc = L.adaqp_spmm_csr_seg_f32(
        graph.indptr.data_ptr(), seg_start, seg_end, graph.indices.data_ptr(), x_local.data_ptr(),
        x_local.stride(0), graph.n_inner, x_halo.data_ptr() if x_halo is not None else None,
        x_halo.stride(0) if x_halo is not None else 0,
        pre.data_ptr() if pre is not None else None, post.data_ptr() if post is not None else None,
        1 if mean else 0, 1 if add_self else 0, accumulate, int(row_begin), row_end, F, out.data_ptr(),
        out.stride(0), _lib.stream_ptr(stream))
    _lib.check(rc, "adaqp_spmm_csr_seg_f32")
    return out
