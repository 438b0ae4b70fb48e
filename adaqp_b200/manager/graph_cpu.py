"""CPU stand-in of LocalGraph for the gloo plumbing configuration (ADAQP_DEVICE=cpu,
BASELINE.json configs[0]: "Reddit GCN 2-part Vanilla on CPU/gloo, no GPU").

NOT the product path and never selected implicitly: it exists so that the control plane,
index contracts and Trainer loop can be exercised by world_size-2 gloo tests on machines
without a GPU.  Aggregation uses torch.sparse on the host.
"""
from __future__ import annotations

import numpy as np
import torch


class CpuGraph(object):
    def __init__(self, indptr, indices, in_degrees, out_degrees, n_inner, n_halo):
        self.device = torch.device("cpu")
        self.n_inner, self.n_halo = int(n_inner), int(n_halo)
        self.nnz = int(len(indices))
        self.csr = torch.sparse_csr_tensor(torch.from_numpy(np.asarray(indptr, np.int64)),
                                           torch.from_numpy(np.asarray(indices, np.int64)),
                                           torch.ones(self.nnz), size=(self.n_inner, self.n_inner + self.n_halo))
        self.local_deg = torch.from_numpy(np.diff(indptr).astype(np.float32))
        self.ndata = {"in_degrees": torch.from_numpy(np.asarray(in_degrees)),
                      "out_degrees": torch.from_numpy(np.asarray(out_degrees))}
        ind = self.ndata["in_degrees"].float().clamp(min=1)
        outd = self.ndata["out_degrees"].float().clamp(min=1)
        self.norm = {"in_-0.5": ind.pow(-0.5), "out_-0.5": outd.pow(-0.5), "out_-1": outd.pow(-1),
                     "in_+1_-1": (ind + 1).pow(-1), "out_+1_-1": (outd + 1).pow(-1)}

    def num_nodes(self):
        return self.n_inner + self.n_halo


def spmm_cpu(graph: CpuGraph, x_local, x_halo, pre, post, mean=False, add_self=False,
             row_begin=0, row_end=None):
    row_end = graph.n_inner if row_end is None else row_end
    x = x_local if x_halo is None or x_halo.shape[0] == 0 else torch.cat([x_local, x_halo], 0)
    if x.shape[0] < graph.n_inner + graph.n_halo:      # central rows never touch halo columns
        x = torch.cat([x, x.new_zeros(graph.n_inner + graph.n_halo - x.shape[0], x.shape[1])], 0)
    xs = x * pre.view(-1, 1) if pre is not None else x
    out = torch.sparse.mm(graph.csr, xs)
    if add_self:
        out = out + xs[:graph.n_inner]
    if mean:
        out = out / graph.local_deg.clamp(min=1).view(-1, 1)
    if post is not None:
        out = out * post[:graph.n_inner].view(-1, 1)
    return out[row_begin:row_end]
