"""Adaptive bit-width assignment end to end on the control plane (2 gloo ranks, CPU): alpha-beta
profiling, traced-variance grouping, gather -> exact solve on rank 0 -> scatter, expansion to rows,
buffer re-generation; plus the on-disk partition path (save_rank_layout / load_rank_layout)."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank), "ADAQP_DEVICE": "cpu", "ADAQP_SYNTH_SCALE": "0.002", "OMP_NUM_THREADS": "1", "ADAQP_SYNTHETIC": "1"})
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    from adaqp_b200.assigner import Assigner
    from adaqp_b200.communicator import Communicator
    from adaqp_b200.helper import BitType, DistGNNType
    from adaqp_b200.manager import GraphEngine
    from adaqp_b200.manager.graphEngine import load_rank_layout, save_rank_layout
    comm = Communicator("gloo", "env://")
    eng = GraphEngine(5, f"{tmp}/parts", "ogbn-products", "quant", DistGNNType.DistGCN, use_parallel=False)
    # on-disk round trip of the prepared layout
    save_rank_layout(eng.layout, f"{tmp}/parts", "ogbn-products")
    comm.barrier()
    again = load_rank_layout(f"{tmp}/parts", "ogbn-products", DistGNNType.DistGCN)
    assert again.n_inner == eng.layout.n_inner and np.array_equal(again.indices, eng.layout.indices)
    assert np.array_equal(again.total_send_idx, eng.layout.total_send_idx)
    dims = [100, 256, 256]
    comm.init_buffer(dims, eng.send_idx, eng.recv_idx, BitType.QUANT, total_send_idx=eng.total_send_idx, num_remote=eng.num_remove)
    asg = Assigner(100, 256, 3, 6, "adaptive", 8, eng.scores, group_size=50, coe_lambda=0.5, assign_cycle=10)
    assert asg.is_tracing and set(asg.cost_model) == {f"{rank}_{p}" for p in eng.send_idx}
    first = asg.get_assignment(eng.send_idx, runtime_scheme="uniform")
    assert all(torch.all(b == 8) for per in first.values() for b in per.values())
    comm.update_buffer(first)
    # pretend one assignment cycle of tracing happened
    n_send = int(eng.total_send_idx.numel())
    g = torch.Generator().manual_seed(rank)
    for key in asg.traced_layer_data:
        asg.traced_layer_data[key] = torch.rand(n_send, generator=g) * (10.0 if "forward" in key else 0.01)
    bits = asg.get_assignment(eng.send_idx)
    assert set(bits) == {"forward0", "forward1", "forward2", "backward1", "backward2"}
    counts = {2: 0, 4: 0, 8: 0}
    for key, per in bits.items():
        assert set(per) == set(eng.send_idx)
        for p, b in per.items():
            lo, hi = eng.send_idx[p]
            assert b.dtype == torch.int32 and b.numel() == hi - lo
            vals = set(b.tolist())
            assert vals <= {2, 4, 8}
            for v in vals:
                counts[v] += int((b == v).sum())
    comm.update_buffer(bits)                       # buffer.py:255-264 with the new layout
    aux = comm.comm_buffer.get_auxillary_buffer("forward1")
    assert set(aux[2]) == set(eng.send_idx) and set(aux[0]) == set(eng.recv_idx)
    assert all(v == 0.0 for v in asg.traced_layer_data.values())   # traces reset for the next cycle
    out.put((rank, counts))


def test_adaptive_assignment_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = sorted(out.get(timeout=5) for _ in procs)
    assert sum(sum(c.values()) for _, c in res) > 0
