"""Per-rank hot-path layout assembled from a raw partition (host side, one-off).

Pipeline of the reference's GraphEngine.__init__ (AdaQP/manager/graphEngine.py:54-76):
convert_partition -> get_send_recv_idx_scores -> reorder_graph -> convert_send_idx
(-> decompose_graph), restated DGL-free on numpy CSR.  `prepare_rank` is the multi-process
form (collectives supplied by the caller), `prepare_all_in_process` wires W ranks inside
one process for tests, smoke() and the single-GPU loopback bench.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple

import numpy as np

from ..helper import DistGNNType
from . import conversion as cv
from .partition_synth import RawPartition, SynthSpec, attach_global_degrees, block_starts, build_raw_partition


@dataclass
class RankLayout:
    rank: int
    world_size: int
    n_central: int
    n_marginal: int
    n_inner: int
    n_halo: int
    indptr: np.ndarray                      # int64 [n_inner + 1], rows = inner nodes [central | marginal]
    indices: np.ndarray                     # int32, source local ids (>= n_inner: halo)
    in_degrees: np.ndarray                  # global, all local nodes
    out_degrees: np.ndarray
    feat: np.ndarray
    label: np.ndarray
    train_mask: np.ndarray
    val_mask: np.ndarray
    test_mask: np.ndarray
    send_idx: Dict[int, Tuple[int, int]]    # peer -> (lo, hi) into total_send_idx
    total_send_idx: np.ndarray              # int64 local inner row ids
    recv_idx: Dict[int, np.ndarray]         # peer -> positions inside the halo block
    scores: Dict[int, Tuple[np.ndarray, np.ndarray]]   # peer -> (forward, backward) aggregation scores
    src_marginal_idx: np.ndarray
    src_central_idx: np.ndarray
    is_bidirected: bool = True


def _finish(raw: RawPartition, recv_idx, send_ids, scores) -> RankLayout:
    ro = cv.reorder_partition(raw, send_ids)
    send_idx, total = cv.convert_send_idx(ro.send_idx)
    sm, sc = cv.decomposition_indices(ro.indptr, ro.indices, ro.n_central, ro.n_inner)
    return RankLayout(rank=raw.rank, world_size=raw.num_parts, n_central=ro.n_central,
                      n_marginal=ro.n_marginal, n_inner=ro.n_inner, n_halo=ro.n_halo,
                      indptr=ro.indptr, indices=ro.indices, in_degrees=ro.in_degrees,
                      out_degrees=ro.out_degrees, feat=ro.feat, label=ro.label,
                      train_mask=ro.train_mask, val_mask=ro.val_mask, test_mask=ro.test_mask,
                      send_idx=send_idx, total_send_idx=total, recv_idx=recv_idx, scores=scores,
                      src_marginal_idx=sm, src_central_idx=sc,
                      is_bidirected=bool(np.array_equal(ro.in_degrees, ro.out_degrees)))


def prepare_rank(spec: SynthSpec, rank: int, model_type: DistGNNType,
                 all_gather: Callable[[object], List[object]]) -> RankLayout:
    """Multi-process form: `all_gather(obj)` returns every rank's obj in rank order."""
    raw = build_raw_partition(spec, rank)
    degs = all_gather(raw.inner_degrees)
    attach_global_degrees(raw, degs, block_starts(spec))
    recv_idx, requests = cv.halo_requests(raw, model_type)
    all_requests = all_gather(requests)
    send_ids, scores = cv.send_side(rank, all_requests)
    return _finish(raw, recv_idx, send_ids, scores)


def prepare_all_in_process(spec: SynthSpec, model_type: DistGNNType = DistGNNType.DistGCN) -> List[RankLayout]:
    W = spec.num_parts
    raws = [build_raw_partition(spec, r) for r in range(W)]
    degs = [r.inner_degrees for r in raws]
    starts = block_starts(spec)
    for r in raws:
        attach_global_degrees(r, degs, starts)
    rr = [cv.halo_requests(r, model_type) for r in raws]
    all_requests = [x[1] for x in rr]
    out = []
    for r in range(W):
        send_ids, scores = cv.send_side(r, all_requests)
        out.append(_finish(raws[r], rr[r][0], send_ids, scores))
    return out
