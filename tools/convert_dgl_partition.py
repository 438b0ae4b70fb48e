"""Convert dgl.distributed.partition_graph output (+ the reference's global degree files) into the
per-rank layout files this package reads (`<part_dir>/<dataset>/<W>part/part<rank>.npz`).

Run it where DGL is installed (it is NOT available in the build image, so this script is
untested here; it restates AdaQP/manager/conversion.py:17-54 and processing.py:40-79 through the
DGL-free functions of adaqp_b200.manager).  All ranks are converted in one process:

    python tools/convert_dgl_partition.py --part-dir data/part_data --dataset reddit --num-parts 4 --model gcn
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def raw_from_dgl(part_config: str, rank: int, world: int, degree_dir: str):
    import dgl
    import torch
    from adaqp_b200.manager.partition_synth import RawPartition
    g, nfeat, _, gpb, _, ntypes, _ = dgl.distributed.load_partition(part_config, rank)
    nt = ntypes[0]
    inner = g.ndata["inner_node"].bool().numpy()
    n_in = int(inner.sum())
    assert inner[:n_in].all(), "DGL stores inner nodes first"
    gid = g.ndata[dgl.NID].numpy()
    u, v = (t.numpy() for t in g.edges())
    keep = v < n_in                                        # in-edges of inner nodes
    import scipy.sparse as sp
    A = sp.coo_matrix((np.ones(keep.sum(), np.int8), (v[keep], u[keep])), shape=(n_in, g.num_nodes())).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    in_deg = torch.load(f"{degree_dir}/in_degrees.pt").numpy()
    out_deg = torch.load(f"{degree_dir}/out_degrees.pt").numpy()
    orig = g.ndata["orig_id"].numpy()
    starts = np.array([gpb.partid2nids(i)[0].item() for i in range(world)] + [gpb._num_nodes()], np.int64)
    return RawPartition(
        rank=rank, num_parts=world, n_inner=n_in, inner_start=int(starts[rank]), starts=starts,
        indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32), halo_gid=gid[n_in:].astype(np.int64),
        halo_part=g.ndata["part_id"].numpy()[n_in:].astype(np.int32), feat=nfeat[nt + "/feat"].numpy().astype(np.float32),
        label=nfeat[nt + "/label"].numpy(), train_mask=nfeat[nt + "/train_mask"].bool().numpy(),
        val_mask=nfeat[nt + "/val_mask"].bool().numpy(), test_mask=nfeat[nt + "/test_mask"].bool().numpy(),
        in_degrees=in_deg[orig], out_degrees=out_deg[orig])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--part-dir", default="data/part_data")
    ap.add_argument("--dataset", required=True)
    ap.add_argument("--num-parts", type=int, required=True)
    ap.add_argument("--model", default="gcn", choices=["gcn", "sage"])
    ap.add_argument("--degree-dir", default=None)
    a = ap.parse_args()
    from adaqp_b200.helper import DistGNNType
    from adaqp_b200.manager import conversion as cv
    from adaqp_b200.manager.graphEngine import save_rank_layout
    from adaqp_b200.manager.layout import _finish
    model = DistGNNType.DistGCN if a.model == "gcn" else DistGNNType.DistSAGE
    cfg = f"{a.part_dir}/{a.dataset}/{a.num_parts}part/{a.dataset}.json"
    deg = a.degree_dir or f"graph_degrees/{a.dataset}"
    raws = [raw_from_dgl(cfg, r, a.num_parts, deg) for r in range(a.num_parts)]
    # DGL does not sort halo nodes by global id; the contract only needs recv positions per owner
    rr = [cv.halo_requests(r, model) for r in raws]
    for r in range(a.num_parts):
        send_ids, scores = cv.send_side(r, [x[1] for x in rr])
        print(save_rank_layout(_finish(raws[r], rr[r][0], send_ids, scores), a.part_dir, a.dataset))


if __name__ == "__main__":
    main()
