"""alpha-beta cost model of every (src, dst) channel: time_ms = alpha * MB + beta.

Reference: AdaQP/assigner/profile.py:18-106 times 200 `dist.send` payloads per peer pair
over gloo and fits a line with np.polyfit.  Here the payload sizes and the fit are the
same, but on the 'p2p' transport the timed operation is the thing the exchange actually
does: a device-to-peer-slab copy over NVLink (CUDA events), so that `adaptive` optimises
the transport in use.  On the 'gloo' transport the reference's send/recv timing is kept.
"""
from __future__ import annotations

import time
from typing import Dict, List, Tuple

import numpy as np
import torch

from ..communicator import BITS_SET
from ..communicator import Communicator as comm
from ..helper import MessageType
from ..manager import GraphEngine as engine


def payload_sizes(num_nodes: int, feat_dim: int, hidden_dim: int, num_data: int) -> np.ndarray:
    """Byte sizes between the all-2-bit and all-8-bit payload of a channel (profile.py:25-28)."""
    low = round(num_nodes * min(feat_dim, hidden_dim) * BITS_SET[0] / 8)
    high = round(num_nodes * max(feat_dim, hidden_dim) * BITS_SET[-1] / 8)
    tol = round(low / 2)
    return np.linspace(max(low - tol, 1), high + tol, num_data).astype(np.int64)


def fit_cost_model(dataset: Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    sizes_mb, times_ms = dataset
    return {k: np.polyfit(np.asarray(sizes_mb[k], np.float64), np.asarray(times_ms[k], np.float64), 1)
            for k in sizes_mb}


def _profile_p2p(feat_dim, hidden_dim, num_data, warmup):
    ex = comm.ctx.comm_buffer.p2p
    rank = comm.get_rank()
    sizes_mb, times_ms = {}, {}
    dev = comm.ctx.device
    key = ex.keys[-1]
    for p, (lo, hi) in ex.send_idx.items():
        sizes = payload_sizes(hi - lo, feat_dim, hidden_dim, num_data)
        cap = ex.layouts[p].halo_off[key]                     # write into the peer's halo block of the last key
        room = 4 * ex.dims[key] * max(ex.layouts[p].num_remote, 1)
        src = torch.zeros(int(min(sizes.max(), room)), dtype=torch.uint8, device=dev)
        dst = _peer_view(ex, p, cap, src.numel())
        ts = []
        for n in sizes:
            n = int(min(n, src.numel()))
            for _ in range(warmup):
                dst[:n].copy_(src[:n])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2 * warmup):
                dst[:n].copy_(src[:n])
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) / (2 * warmup))
        sizes_mb[f"{rank}_{p}"] = np.minimum(sizes, src.numel()) / (1024 ** 2)
        times_ms[f"{rank}_{p}"] = np.asarray(ts)
    comm.barrier()
    return sizes_mb, times_ms


def _peer_view(ex, p: int, offset: int, nbytes: int) -> torch.Tensor:
    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                  "data": (ex.peer_base[p] + offset, False), "version": 2}
    return torch.as_tensor(h, device=ex.device)


def _profile_gloo(feat_dim, hidden_dim, num_data, warmup):
    """profile.py:46-95: sender times dist.send of every payload, receivers drain."""
    rank, W = comm.get_rank(), comm.get_world_size()
    send_idx, recv_idx = engine.ctx.send_idx, engine.ctx.recv_idx
    sizes_mb, times_ms = {}, {}
    for sender in range(W):
        if sender == rank:
            for p, (lo, hi) in send_idx.items():
                sizes = payload_sizes(hi - lo, feat_dim, hidden_dim, num_data)
                ts = []
                for n in sizes:
                    buf = torch.zeros(int(n), dtype=torch.uint8)
                    reps = []
                    for it in range(1, 3 * warmup):
                        t0 = time.time()
                        comm.sync_send(buf, p, MessageType.DATA)
                        if it > warmup:
                            reps.append(time.time() - t0)
                    ts.append(1000 * sum(reps) / max(len(reps), 1))
                sizes_mb[f"{rank}_{p}"] = sizes / (1024 ** 2)
                times_ms[f"{rank}_{p}"] = np.asarray(ts)
        elif sender in recv_idx:
            for n in payload_sizes(len(recv_idx[sender]), feat_dim, hidden_dim, num_data):
                buf = torch.zeros(int(n), dtype=torch.uint8)
                for _ in range(1, 3 * warmup):
                    comm.sync_recv(buf, sender, MessageType.DATA)
        comm.barrier()
    return sizes_mb, times_ms


def generate_cost_model_dataset(feat_dim: int, hidden_dim: int, num_data: int, warmup: int):
    if comm.ctx.transport == "p2p":
        return _profile_p2p(feat_dim, hidden_dim, num_data, max(warmup, 1))
    return _profile_gloo(feat_dim, hidden_dim, num_data, max(warmup, 1))
