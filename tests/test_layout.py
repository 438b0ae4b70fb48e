"""Host-side layout contract on synthetic partitions: no GPU."""
import numpy as np
import pytest

from adaqp_b200.helper import DistGNNType
from adaqp_b200.manager.layout import prepare_all_in_process
from adaqp_b200.manager.partition_synth import SynthSpec, block_starts, build_all_partitions


def small_spec(W=4, n=2000, deg=12, F=20, chi=0.25, seed=3):
    return SynthSpec(name="t", num_nodes=n, num_edges=n * deg, num_parts=W, num_feats=F, num_classes=5,
                     cross_fraction=chi, community_size=64, seed=seed)


def global_adjacency(raws, spec):
    """Rebuild the global directed edge set from the per-rank partitions."""
    edges = set()
    for r in raws:
        gid = np.concatenate([np.arange(r.n_inner) + r.inner_start, r.halo_gid])
        dst = np.repeat(np.arange(r.n_inner), np.diff(r.indptr)) + r.inner_start
        src = gid[r.indices]
        edges |= set(zip(src.tolist(), dst.tolist()))
    return edges


@pytest.mark.parametrize("W", [1, 2, 4])
def test_partitions_are_consistent(W):
    spec = small_spec(W=W)
    raws = build_all_partitions(spec)
    E = global_adjacency(raws, spec)
    assert all((v, u) in E for (u, v) in E)                     # symmetric
    assert all((i, i) in E for i in range(spec.num_nodes))       # self-loops
    deg = np.zeros(spec.num_nodes, np.int64)
    for (u, v) in E:
        deg[v] += 1
    for r in raws:
        gid = np.concatenate([np.arange(r.n_inner) + r.inner_start, r.halo_gid])
        np.testing.assert_array_equal(r.in_degrees, deg[gid])    # global degrees incl. halo
        assert np.all(np.diff(r.halo_gid) > 0)
        if W == 1:
            assert r.n_halo == 0


@pytest.mark.parametrize("model", [DistGNNType.DistGCN, DistGNNType.DistSAGE])
def test_rank_layout_contract(model):
    spec = small_spec(W=4)
    raws = build_all_partitions(spec)
    lays = prepare_all_in_process(spec, model)
    starts = block_starts(spec)
    for L, raw in zip(lays, raws):
        assert L.n_central + L.n_marginal == L.n_inner
        dst = np.repeat(np.arange(L.n_inner), np.diff(L.indptr))
        has_halo = np.zeros(L.n_inner, bool)
        has_halo[dst[L.indices >= L.n_inner]] = True
        assert not has_halo[:L.n_central].any() and has_halo[L.n_central:].all()
        # recv_idx partitions the halo block; send rows correspond 1:1 to the peer's halo rows
        allpos = np.sort(np.concatenate(list(L.recv_idx.values())))
        np.testing.assert_array_equal(allpos, np.arange(L.n_halo))
        assert list(L.send_idx) == sorted(L.send_idx)
        for p, (lo, hi) in L.send_idx.items():
            peer = lays[p]
            assert hi - lo == peer.recv_idx[L.rank].size
            # row j sent to p is the node p sees at halo position recv_idx[me][j]
            new_to_gid = np.empty(L.n_inner, np.int64)
            new_to_gid[np.arange(L.n_inner)] = 0
            inv = np.empty(L.n_inner, np.int64)
            # invert the reorder through features: compare feature rows instead of ids
            mine = L.feat[L.total_send_idx[lo:hi]]
            want_gid = raws[p].halo_gid[peer.recv_idx[L.rank]]
            np.testing.assert_array_equal(mine, raws[L.rank].feat[want_gid - starts[L.rank]])
            assert L.scores[p][0].shape == (hi - lo,) and np.all(L.scores[p][0] > 0)
        # decomposition helper indices
        cols_c = L.indices[:L.indptr[L.n_central]]
        assert set(L.src_marginal_idx.tolist()) == set(cols_c[cols_c >= L.n_central].tolist())
        # reorder preserved the multiset of (feature row, degree)
        assert np.isclose(L.feat.sum(), raw.feat.sum(), rtol=1e-4)
        np.testing.assert_array_equal(np.sort(L.in_degrees[:L.n_inner]), np.sort(raw.in_degrees[:raw.n_inner]))


def test_directed_partition_is_detected_and_refused():
    """is_bidirected follows conversion.py:28-32 (all(in_degrees == out_degrees)); a directed layout is refused before any
    device work (the reference's reversed backward graph cannot run either)."""
    from adaqp_b200.manager import GraphEngine
    from adaqp_b200.manager import layout as lay
    from adaqp_b200.manager import conversion as cv
    spec = small_spec(W=1, n=300)
    L = prepare_all_in_process(spec)[0]
    assert L.is_bidirected
    raws = build_all_partitions(spec)
    raws[0].out_degrees = raws[0].out_degrees.copy()
    raws[0].out_degrees[0] += 1                                   # one node with out-degree != in-degree
    recv_idx, requests = cv.halo_requests(raws[0], DistGNNType.DistGCN)
    send_ids, scores = cv.send_side(0, [requests])
    Ld = lay._finish(raws[0], recv_idx, send_ids, scores)
    assert not Ld.is_bidirected
    with pytest.raises(NotImplementedError, match="directed"):
        GraphEngine(1, None, "t", "full", DistGNNType.DistGCN, layout=Ld)
