"""oracle/ref_path.py (the reference-flow arm) on CPU: its aggregation restatements (ops.py:17-67 with torch.sparse in place of
DGL's update_all) against the C oracle's float64 aggregation, and its model factories.  The flow itself (gloo ring, pinned
staging, reference kernels) needs CUDA and is exercised by bench.py --impl reference and tools/parity_check.py on the GPU."""
import numpy as np
import pytest
import torch

from adaqp_b200.helper import DistGNNType
from adaqp_b200.manager.layout import prepare_all_in_process
from adaqp_b200.manager.partition_synth import SynthSpec
from oracle import oracle as O
from oracle import ref_path


def _state(L, kind):
    """A RefState with only what the aggregation methods read (its constructor needs CUDA streams)."""
    st = object.__new__(ref_path.RefState)
    ind = torch.from_numpy(L.in_degrees).float().clamp(min=1)
    outd = torch.from_numpy(L.out_degrees).float().clamp(min=1)
    st.norm = {"in": ind.pow(-0.5), "out": outd.pow(-0.5), "out_-1": torch.pow(outd, -1)}
    st.local_indeg = torch.from_numpy(np.diff(L.indptr).astype(np.float32))
    st.kind = kind
    n_all = L.n_inner + L.n_halo
    st.full = torch.sparse_csr_tensor(torch.from_numpy(L.indptr.astype(np.int64)), torch.from_numpy(L.indices.astype(np.int64)),
                                      torch.ones(len(L.indices)), size=(L.n_inner, n_all))
    return st


@pytest.mark.parametrize("kind", ["gcn", "sage"])
@pytest.mark.parametrize("backward", [False, True])
def test_reference_flow_aggregation_matches_oracle(kind, backward):
    spec = SynthSpec(name="t", num_nodes=1500, num_edges=1500 * 12, num_parts=3, num_feats=24, num_classes=5,
                     cross_fraction=0.3, community_size=64, seed=5)
    L = prepare_all_in_process(spec, DistGNNType.DistGCN if kind == "gcn" else DistGNNType.DistSAGE)[1]
    st = _state(L, kind)
    rng = np.random.RandomState(3)
    x = rng.standard_normal((L.n_inner + L.n_halo, 24)).astype(np.float32)
    got = st.agg(st.full, torch.from_numpy(x), 0, L.n_inner, backward).numpy()
    fn = O.gcn_aggregation if kind == "gcn" else O.sage_aggregation
    want = fn(L.indptr, L.indices.astype(np.int64), x, L.in_degrees, L.out_degrees, L.n_inner, backward=backward)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # rows shorter than the column space are zero-extended (central rows never touch halo columns)
    got2 = st.agg(st.full, torch.from_numpy(np.concatenate([x[:L.n_inner], np.zeros_like(x[L.n_inner:])])), 0, L.n_inner, backward)
    got3 = st.agg(st.full, torch.from_numpy(x[:L.n_inner].copy()), 0, L.n_inner, backward)
    assert torch.equal(got2, got3)


def test_model_factories_mirror_the_reference_modules():
    g = ref_path.make_model("gcn", [20, 16, 16], 5, 0.5)
    s = ref_path.make_model("sage", [20, 16, 16], 5, 0.5)
    assert isinstance(g, ref_path.RefGCN) and isinstance(s, ref_path.RefSAGE)
    assert [tuple(w.shape) for w in g.w] == [(20, 16), (16, 16), (16, 5)] and len(g.norms) == 2
    assert sum(p.numel() for p in s.parameters()) > sum(p.numel() for p in g.parameters())    # fc_self + fc_neigh
