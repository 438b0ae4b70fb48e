"""Distributed aggregation ops (autograd Functions) over the fused exchange + CSR SpMM.

Mirror of AdaQP/model/ops.py: `GCN_aggregation` / `SAGE_aggregation` (:17-67),
`DistAggConv` / `DistAggSAGE` (:69-111), `full_graph_propagation` (:132-154) and
`decomposed_graph_propagation` (:156-193) keep names, arguments and results.

B200 design: the aggregation reads local rows and halo rows in place (no torch.cat), the
central / marginal split is a row range of one CSR (no copy buffers, no host sync), and
in the decomposed path the exchange kernels run on the side stream while the central rows
aggregate on the default stream; ordering is by CUDA events only.
"""
from __future__ import annotations

from typing import Any, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from ..communicator import Communicator as comm
from ..helper import BitType, ProprogationMode
from ..manager import DecompGraph
from ..manager import GraphEngine as engine
from ..manager.graph import LocalGraph, spmm
from ..manager.graphEngine import RowRange
from .op_util import halo_exchange, msg_all2all_GLOO


def _split(graph, feats: Tensor, x_halo: Tensor = None):
    g = graph.graph if isinstance(graph, RowRange) else graph
    if x_halo is None and feats.shape[0] > g.n_inner:
        return g, feats[:g.n_inner], feats[g.n_inner:]
    return g, feats, x_halo


def _run(g, x_local, x_halo, pre, post, mean, add_self, lo, hi, out=None, part=None):
    if isinstance(g, LocalGraph):
        return spmm(g, x_local, x_halo, pre, post, mean=mean, add_self=add_self, row_begin=lo, row_end=hi, out=out,
                    part=part)
    from ..manager.graph_cpu import spmm_cpu          # gloo plumbing mode
    res = spmm_cpu(g, x_local, x_halo, pre, post, mean=mean, add_self=add_self, row_begin=lo, row_end=hi)
    if out is not None:
        out.copy_(res)
        return out
    return res


def GCN_aggregation(graph, feats: Tensor, mode: ProprogationMode = ProprogationMode.Forward,
                    x_halo: Tensor = None, out: Tensor = None, part: str = None) -> Tensor:
    """out[v] = norm2[v] * sum_{u->v} norm1[u] x[u] with global-degree norms (ops.py:17-32).
    `feats` may be cat(local, halo) as in the reference, or the local rows with `x_halo`."""
    g, x_local, x_halo = _split(graph, feats, x_halo)
    lo, hi = (graph.begin, graph.end) if isinstance(graph, RowRange) else (0, g.n_inner)
    if mode == ProprogationMode.Forward:
        pre, post = g.norm["out_-0.5"], g.norm["in_-0.5"]
    elif mode == ProprogationMode.Backward:
        pre, post = g.norm["in_-0.5"], g.norm["out_-0.5"]
    else:
        raise ValueError(f"Invalid mode {mode}")
    return _run(g, x_local, x_halo, pre, post, False, False, lo, hi, out, part)


def SAGE_aggregation(graph, feats: Tensor, mode: ProprogationMode = ProprogationMode.Forward,
                     aggregator_type="mean", x_halo: Tensor = None, out: Tensor = None, part: str = None) -> Tensor:
    """ops.py:34-67: 'mean' = mean over in-neighbours (fwd) / sum of x[u]/outdeg[u] (bwd);
    'gcn' = (sum + self) / (indeg + 1) (fwd) / sum + self of x/(outdeg+1) (bwd)."""
    g, x_local, x_halo = _split(graph, feats, x_halo)
    lo, hi = (graph.begin, graph.end) if isinstance(graph, RowRange) else (0, g.n_inner)
    if mode == ProprogationMode.Forward:
        if aggregator_type == "mean":
            return _run(g, x_local, x_halo, None, None, True, False, lo, hi, out, part)
        if aggregator_type == "gcn":
            return _run(g, x_local, x_halo, None, g.norm["in_+1_-1"], False, True, lo, hi, out, part)
    elif mode == ProprogationMode.Backward:
        if aggregator_type == "mean":
            return _run(g, x_local, x_halo, g.norm["out_-1"], None, False, False, lo, hi, out, part)
        if aggregator_type == "gcn":
            return _run(g, x_local, x_halo, g.norm["out_+1_-1"], None, False, True, lo, hi, out, part)
    else:
        raise ValueError(f"Invalid mode {mode}")
    raise ValueError(f"Invalid aggregator_type {aggregator_type}")


def _aggregate(class_name: str, graph, x_local, x_halo, mode, out, part=None):
    if class_name == "DistAggConv":
        return GCN_aggregation(graph, x_local, mode=mode, x_halo=x_halo, out=out, part=part)
    if class_name == "DistAggSAGE":
        return SAGE_aggregation(graph, x_local, mode=mode, aggregator_type=engine.ctx.agg_type, x_halo=x_halo, out=out,
                                part=part)
    raise ValueError(f"Invalid class_name {class_name}")


def _eval_layer0(class_name: str, ctx, local_messages: Tensor, graph, layer: int, is_train: bool):
    """every epoch's evaluation forward (trainer.py:181) exchanges and aggregates the
    CONSTANT input features of layer 0 in fp32 -- the result never changes, so it is computed once
    and reused (keyed on the feature tensor's storage and version).  Off while the Assigner traces
    (eval passes feed its variance statistics, op_util.py:91-99) or with ADAQP_EVAL_CACHE=0."""
    import os
    from ..assigner import Assigner as assigner
    eng = engine.ctx
    fn = decomposed_graph_propagation if eng.use_parallel else full_graph_propagation
    usable = (not is_train and layer == 0 and os.environ.get("ADAQP_EVAL_CACHE", "1") != "0"
              and not (assigner.ctx is not None and assigner.ctx.is_tracing))
    if not usable:
        return fn(ctx, local_messages, graph, layer, is_train, ProprogationMode.Forward, class_name)
    # keyed on the tensor OBJECT (held alive by the cache, so its address cannot be recycled), its version
    # counter, the graph and the aggregator; only the engine's own constant feature matrix is cached, so
    # every rank takes the same branch (a hit skips the exchange: ranks must agree or the key's sequence
    # numbers would diverge)
    if local_messages is not eng.feats:
        return fn(ctx, local_messages, graph, layer, is_train, ProprogationMode.Forward, class_name)
    key = (class_name, local_messages._version, tuple(local_messages.shape), id(graph),
           eng._agg_type if class_name == "DistAggSAGE" else None)
    cache = getattr(eng, "_eval_layer0_cache", None)
    if cache is None or cache[0] is not local_messages or cache[1] != key:
        out = fn(ctx, local_messages, graph, layer, is_train, ProprogationMode.Forward, class_name)
        eng._eval_layer0_cache = (local_messages, key, out)
        return out
    ctx.saved = layer
    return cache[2]


class DistAggConv(Function):
    """Aggregation of local + remote neighbours for GCN (ops.py:69-89)."""

    @staticmethod
    def forward(ctx, local_messages: Tensor, graph, layer: int, is_train: bool) -> Tensor:
        return _eval_layer0(DistAggConv.__name__, ctx, local_messages, graph, layer, is_train)

    @staticmethod
    def backward(ctx: Any, *grad_outputs: Tuple[Tensor, ...]):
        fn = decomposed_graph_propagation if engine.ctx.use_parallel else full_graph_propagation
        return fn(ctx, grad_outputs[0].contiguous(), engine.ctx.bwd_graph, ctx.saved, True,
                  ProprogationMode.Backward, DistAggConv.__name__)


class DistAggSAGE(Function):
    """Aggregation of local + remote neighbours for GraphSAGE (ops.py:91-111)."""

    @staticmethod
    def forward(ctx, local_messages: Tensor, graph, layer: int, is_train: bool) -> Tensor:
        return _eval_layer0(DistAggSAGE.__name__, ctx, local_messages, graph, layer, is_train)

    @staticmethod
    def backward(ctx: Any, *grad_outputs: Tuple[Tensor, ...]):
        fn = decomposed_graph_propagation if engine.ctx.use_parallel else full_graph_propagation
        return fn(ctx, grad_outputs[0].contiguous(), engine.ctx.bwd_graph, ctx.saved, True,
                  ProprogationMode.Backward, DistAggSAGE.__name__)


_SPLIT = None


def _split_marginal() -> bool:
    """Two-pass marginal aggregation (default; ADAQP_MARGINAL_SPLIT=0 restores the reference's
    single pass): local sources overlap the exchange, the halo sources accumulate afterwards.
    Removes the exposed wait at the price of a second pass over the marginal rows; measured
    both ways in profiles/r01_overlap.md."""
    global _SPLIT
    if _SPLIT is None:
        import os
        _SPLIT = os.environ.get("ADAQP_MARGINAL_SPLIT", "1") != "0"
    return _SPLIT


def _finish(ctx, out: Tensor, layer: int, mode: ProprogationMode):
    if mode == ProprogationMode.Forward:
        ctx.saved = layer
        return out
    return out, None, None, None


def full_graph_propagation(ctx, local_messages: Tensor, graph, layer: int, is_train: bool,
                           mode: ProprogationMode, class_name: str):
    """Exchange, then aggregate every inner row (ops.py:132-154)."""
    name = f"forward{layer}" if mode == ProprogationMode.Forward else f"backward{layer}"
    local_messages = local_messages.contiguous()
    timer = engine.ctx.timer
    g = graph.full if isinstance(graph, DecompGraph) else graph
    if comm.ctx.transport == "p2p":
        quant = engine.ctx.bit_type == BitType.QUANT and is_train
        with timer.record_events(f"{name}_quantization" if quant else f"{name}_communication"):
            pend = halo_exchange(local_messages, name, is_train)
        with timer.record_events(f"{name}_full_aggregation"):
            out = _aggregate(class_name, g, local_messages, pend.halo, mode, None)
        pend.release()
    else:
        send_messages = local_messages[engine.ctx.total_send_idx]
        remote = msg_all2all_GLOO(send_messages, name, is_train)
        with timer.record(f"{name}_full_aggregation"):
            out = _aggregate(class_name, g, local_messages, remote, mode, None)
    return _finish(ctx, out, layer, mode)


def decomposed_graph_propagation(ctx, local_messages: Tensor, graph, layer: int, is_train: bool,
                                 mode: ProprogationMode, class_name: str):
    """Exchange on the side stream || central rows on the default stream, then the marginal
    rows once the halo has landed (ops.py:156-193)."""
    assert isinstance(graph, DecompGraph), f"graph must be a DecompGraph, but got {type(graph)}"
    name = f"forward{layer}" if mode == ProprogationMode.Forward else f"backward{layer}"
    local_messages = local_messages.contiguous()
    eng, timer = engine.ctx, engine.ctx.timer
    if comm.ctx.transport != "p2p":
        # gloo plumbing transport: the exchange runs in the helper thread while this thread aggregates the central
        # rows, as in the reference (ops.py:164-177: marginal_pool.apply_async ... response.get())
        send_messages = local_messages[eng.total_send_idx]
        pool = getattr(eng, "marginal_pool", None)
        response = pool.apply_async(msg_all2all_GLOO, (send_messages, name, is_train)) if pool is not None else None
        out = local_messages.new_empty((eng.num_inner, local_messages.shape[1]))
        with timer.record(f"{name}_central_aggregation"):
            _aggregate(class_name, graph.central_graph, local_messages, None, mode, out[:eng.num_central])
        remote = response.get() if response is not None else msg_all2all_GLOO(send_messages, name, is_train)
        with timer.record(f"{name}_marginal_aggregation"):
            _aggregate(class_name, graph.marginal_graph, local_messages, remote, mode, out[eng.num_central:])
        return _finish(ctx, out, layer, mode)
    main, side = torch.cuda.current_stream(), eng.marginal_stream
    quant = eng.bit_type == BitType.QUANT and is_train
    ready = torch.cuda.Event()
    ready.record(main)                       # local_messages is produced on the default stream
    side.wait_event(ready)
    with timer.record_events(f"{name}_quantization" if quant else f"{name}_communication", stream=side):
        pend = halo_exchange(local_messages, name, is_train, stream=side)
    landed = torch.cuda.Event(enable_timing=True)
    landed.record(side)
    out = local_messages.new_empty((eng.num_inner, local_messages.shape[1]))
    with timer.record_events(f"{name}_central_aggregation"):
        _aggregate(class_name, graph.central_graph, local_messages, None, mode, out[:eng.num_central])
    if _split_marginal():
        # the marginal rows' LOCAL-source neighbours do not need the halo either: aggregate them while
        # the exchange is still in flight; only the halo-source segment of each row waits for it
        with timer.record_events(f"{name}_marginal_aggregation_local"):
            _aggregate(class_name, graph.marginal_graph, local_messages, None, mode, out[eng.num_central:], part="local")
        overlappable_done = torch.cuda.Event(enable_timing=True)
        overlappable_done.record(main)
        timer.record_exposed(name, overlappable_done, landed)
        main.wait_event(landed)
        with timer.record_events(f"{name}_marginal_aggregation_halo"):
            _aggregate(class_name, graph.marginal_graph, local_messages, pend.halo, mode, out[eng.num_central:], part="halo")
    else:
        central_done = torch.cuda.Event(enable_timing=True)
        central_done.record(main)
        timer.record_exposed(name, central_done, landed)
        main.wait_event(landed)
        with timer.record_events(f"{name}_marginal_aggregation"):
            _aggregate(class_name, graph.marginal_graph, local_messages, pend.halo, mode, out[eng.num_central:])
    pend.release()
    local_messages.record_stream(side)
    return _finish(ctx, out, layer, mode)
