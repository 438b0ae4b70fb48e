"""the real-partition ingest chain (tools/convert_dgl_partition.py) on a hand-built,
DGL-shaped fixture -- no DGL import.  The fixture is what dgl.distributed.load_partition would hand
back for partitions of the synthetic graph: local COO edge list in arbitrary order, with duplicate
edges and extra edges whose destination is a halo node (DGL keeps those for >1-hop halos; the hot
path must drop them), `inner_node` / NID / part_id / orig_id node data and GLOBAL degree arrays.
The converted layouts must equal the synthetic generator's own, field by field, and survive the
plain-array file format."""
import dataclasses
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from adaqp_b200.helper import DistGNNType
from adaqp_b200.manager.layout import prepare_all_in_process
from adaqp_b200.manager.partition_synth import SynthSpec, build_all_partitions
from tools import convert_dgl_partition as conv


def dgl_shaped(raw, rng, n_total):
    """One RawPartition of the generator -> the arrays DGL's partition loader returns."""
    n_in, n_halo = raw.n_inner, raw.n_halo
    dst = np.repeat(np.arange(n_in, dtype=np.int64), np.diff(raw.indptr))
    src = raw.indices.astype(np.int64)
    # extra edges INTO halo nodes (dropped by the conversion) and duplicated in-edges (collapsed)
    k = max(4, n_halo // 3)
    extra_u = rng.integers(0, n_in, k)
    extra_v = n_in + rng.integers(0, max(n_halo, 1), k) if n_halo else np.zeros(0, np.int64)
    dup = rng.integers(0, src.size, src.size // 10)
    u = np.concatenate([src, extra_u[:extra_v.size], src[dup]])
    v = np.concatenate([dst, extra_v, dst[dup]])
    perm = rng.permutation(u.size)
    nid = np.concatenate([raw.inner_start + np.arange(n_in, dtype=np.int64), raw.halo_gid])
    # global degree arrays indexed by ORIGINAL ids; orig_id is a permutation of the global ids
    orig_of_gid = rng.permutation(n_total)
    indeg_g = np.zeros(n_total, np.int64)
    outdeg_g = np.zeros(n_total, np.int64)
    return {"edges_u": u[perm], "edges_v": v[perm],
            "inner_node": np.concatenate([np.ones(n_in, bool), np.zeros(n_halo, bool)]), "nid": nid,
            "part_id": np.concatenate([np.full(n_in, raw.rank, np.int32), raw.halo_part]),
            "orig_id": orig_of_gid[nid], "starts": raw.starts, "feat": raw.feat, "label": raw.label,
            "train_mask": raw.train_mask, "val_mask": raw.val_mask, "test_mask": raw.test_mask,
            "_orig_of_gid": orig_of_gid, "_deg_slots": (indeg_g, outdeg_g)}


@pytest.mark.parametrize("model", [DistGNNType.DistGCN, DistGNNType.DistSAGE])
@pytest.mark.parametrize("W", [2, 4])
def test_fixture_converts_to_the_generators_layout(W, model, tmp_path):
    spec = SynthSpec(name="fixture", num_nodes=1800, num_edges=1800 * 12, num_parts=W, num_feats=24, num_classes=5,
                     cross_fraction=0.25, community_size=64, seed=3)
    want = prepare_all_in_process(spec, model)
    raws = build_all_partitions(spec)                       # global degrees attached
    rng = np.random.default_rng(0)
    arrays = [dgl_shaped(r, rng, spec.num_nodes) for r in raws]
    # one shared pair of global degree files, as the reference caches them (partition.py:67-68)
    indeg = np.zeros(spec.num_nodes, np.int64)
    outdeg = np.zeros(spec.num_nodes, np.int64)
    orig_of_gid = arrays[0]["_orig_of_gid"]
    for a, r in zip(arrays, raws):
        a["orig_id"] = orig_of_gid[a["nid"]]
        indeg[a["orig_id"]] = r.in_degrees
        outdeg[a["orig_id"]] = r.out_degrees
    for a in arrays:
        a["in_degrees_global"], a["out_degrees_global"] = indeg, outdeg
    got = conv.convert([conv.raw_from_arrays(a, r, W) for r, a in enumerate(arrays)], model)
    from adaqp_b200.manager.graphEngine import read_rank_layout, save_rank_layout
    for g, w in zip(got, want):
        back = read_rank_layout(save_rank_layout(g, str(tmp_path), "fixture"))
        for lay in (g, back):
            for f in dataclasses.fields(w):
                a, b = getattr(lay, f.name), getattr(w, f.name)
                if isinstance(b, dict):
                    assert set(a) == set(b), f.name
                    for k in b:
                        x, y = a[k], b[k]
                        if isinstance(y, tuple) and isinstance(y[0], np.ndarray):
                            assert all(np.array_equal(p, q) for p, q in zip(x, y)), (f.name, k)
                        elif isinstance(y, tuple):
                            assert tuple(int(t) for t in x) == tuple(int(t) for t in y), (f.name, k)
                        else:
                            assert np.array_equal(x, y), (f.name, k)
                elif isinstance(b, np.ndarray):
                    assert np.array_equal(a, b), f.name
                else:
                    assert a == b, f.name


def test_layout_files_hold_no_pickles(tmp_path):
    spec = SynthSpec(name="fixture", num_nodes=600, num_edges=600 * 8, num_parts=2, num_feats=8, num_classes=3,
                     cross_fraction=0.2, community_size=32, seed=1)
    lay = prepare_all_in_process(spec)[0]
    from adaqp_b200.manager.graphEngine import save_rank_layout
    path = save_rank_layout(lay, str(tmp_path), "fixture")
    z = np.load(path, allow_pickle=False)          # would raise on any object array
    assert all(z[k].dtype != object for k in z.files)
