"""DistGCN: aggregate-then-transform GCN layers over the distributed aggregation op
(reference: AdaQP/model/distGCN.py:9-85).  The dense feature x weight product runs on the
tcgen05 tensor cores through adaqp_b200.dense (3xTF32, fp32 result; torch.matmul where the kernel does not apply
or with ADAQP_GEMM=0); LayerNorm, dropout and ReLU stay torch ops."""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn import init
from torch.nn.parameter import Parameter

from .. import dense, fused
from .ops import DistAggConv


class DistGCNConv(nn.Module):
    def __init__(self, in_feats: int, out_feats: int, weight=True, bias: bool = True, activation: Any = None):
        super().__init__()
        self._in_feats, self._out_feats, self._activation = in_feats, out_feats, activation
        self.weight = Parameter(torch.empty(in_feats, out_feats)) if weight else None
        self.bias = Parameter(torch.empty(out_feats)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            init.xavier_uniform_(self.weight)
        if self.bias is not None:
            init.zeros_(self.bias)

    def forward(self, feats: Tensor, graph, layer: int) -> Tensor:
        rst = DistAggConv.apply(feats, graph, layer, self.training)     # exchange + aggregation
        if self.weight is not None:
            rst = dense.linear(rst, self.weight, self.bias)       # rst @ W + b (distGCN.py:45-47)
        elif self.bias is not None:
            rst = rst + self.bias
        return self._activation(rst) if self._activation is not None else rst


class DistGCN(nn.Module):
    def __init__(self, in_feats: int, h_feats: int, num_classes: int, num_layers: int, drop_rate: float,
                 use_norm: bool = True):
        super().__init__()
        dims = [in_feats] + [h_feats] * (num_layers - 1) + [num_classes]
        self.convs = nn.ModuleList(DistGCNConv(dims[i], dims[i + 1]) for i in range(num_layers))
        if use_norm:
            self.norms = nn.ModuleList(nn.LayerNorm(h_feats) for _ in range(num_layers - 1))
        self.drop_rate = drop_rate

    def reset_parameters(self):
        for m in list(self.convs) + list(getattr(self, "norms", [])):
            m.reset_parameters()

    def forward(self, g, feats: Tensor) -> Tensor:
        last = len(self.convs) - 1
        for i in range(last):
            feats = self.convs[i](feats, g, i)
            feats = F.dropout(feats, p=self.drop_rate, training=self.training)
            if hasattr(self, "norms"):
                feats = fused.layer_norm_relu(feats, self.norms[i])      # relu(norms[i](feats)), one pass (csrc/norm.cu)
            else:
                feats = F.relu(feats, inplace=True)
        return self.convs[last](feats, g, last)
