"""The host-built work-item tables of the fused exchange reproduce the reference wire format
(checked against the CPU oracle, no GPU): executing the items with the oracle's codec yields the
byte stream / params of oracle.mixed_quantize, and the receive items scatter to the right halo rows."""
import numpy as np
import pytest

from adaqp_b200.communicator.p2p import build_recv_items, build_send_items, qsize
from oracle import oracle as O


@pytest.mark.parametrize("F,seed", [(13, 0), (100, 1), (256, 2)])
def test_send_items_reproduce_wire_format(F, seed):
    rng = np.random.RandomState(seed)
    n_local = 300
    x = rng.standard_normal((n_local, F)).astype(np.float32)
    send_idx = {1: (0, 37), 3: (37, 37 + 64), 4: (101, 101 + 5)}
    total = rng.randint(0, n_local, 106).astype(np.int64)
    bits = {1: np.array([2, 4, 8])[rng.randint(0, 3, 37)], 3: np.full(64, 4), 4: np.array([8, 2, 2, 8, 8])}
    items, inc = build_send_items(list(send_idx), send_idx, total, bits, F)
    gen_seed, base = 99, 40
    off = base
    for ci, p in enumerate(send_idx):
        lo, hi = send_idx[p]
        rows = x[total[lo:hi]]
        want_q, want_prm, valid, off_after = O.mixed_quantize(rows, bits[p], gen_seed, off)
        got_q = np.zeros(want_q.size, np.uint8)
        got_prm = np.zeros((2, hi - lo), np.uint16)
        mine = items[items["chan"] == ci]
        assert set(np.unique(mine["bits"]).tolist()) == set(np.unique(bits[p]).tolist())
        for b in np.unique(mine["bits"]):
            seg = mine[mine["bits"] == b]
            wpt = 8 // b
            # all items of a segment share one generator call
            assert np.all(seg["rel_offset"] == seg["rel_offset"][0])
            src = seg["src_row"][:, :wpt].reshape(-1)
            src = src[src >= 0]
            pos = seg["send_pos"][:, :wpt].reshape(-1)
            np.testing.assert_array_equal(src, total[pos[pos >= 0]])
            sub = x[src]
            rmin, _, scale = O.minmax_scale(sub, int(b))
            payload = O.pack(sub, rmin, scale, int(b), gen_seed, base + int(seg["rel_offset"][0]))
            for k, it in enumerate(seg):
                assert it["group"] == k and it["nrows"] == min(wpt, src.size - k * wpt)
                got_q[it["dst_off"]:it["dst_off"] + F] = payload[k * F:(k + 1) * F]
                n = int(it["nrows"])
                got_prm[0, it["param_pos"]:it["param_pos"] + n] = O.to_bf16(scale[k * wpt:k * wpt + n])
                got_prm[1, it["param_pos"]:it["param_pos"] + n] = O.to_bf16(rmin[k * wpt:k * wpt + n])
        np.testing.assert_array_equal(got_q[valid], want_q[valid])
        np.testing.assert_array_equal(got_prm, want_prm)
        off = off_after
    assert off - base == inc


def test_recv_items_scatter_to_halo_rows():
    rng = np.random.RandomState(5)
    F = 24
    recv_idx = {0: np.array([3, 4, 5, 9, 10]), 2: np.arange(20, 51)}
    bits = {0: np.array([8, 2, 2, 4, 2]), 2: np.array([2, 4, 8])[rng.randint(0, 3, 31)]}
    items, wire = build_recv_items(list(recv_idx), recv_idx, bits, F)
    for ci, p in enumerate(recv_idx):
        mine = items[items["chan"] == ci]
        expect_bytes = sum(qsize(int((bits[p] == b).sum()), b, F) for b in (2, 4, 8) if (bits[p] == b).any())
        assert wire[p] == (expect_bytes, recv_idx[p].size)
        seen = []
        seg_off = prm = 0
        for b in (2, 4, 8):
            ids = np.nonzero(bits[p] == b)[0]
            if not ids.size:
                continue
            seg = mine[mine["bits"] == b]
            wpt = 8 // b
            np.testing.assert_array_equal(seg["src_off"], seg_off + np.arange(len(seg)) * F)
            np.testing.assert_array_equal(seg["param_pos"], prm + np.arange(len(seg)) * wpt)
            dst = seg["dst_row"][:, :wpt].reshape(-1)
            np.testing.assert_array_equal(dst[dst >= 0], recv_idx[p][ids])
            seen.append(dst[dst >= 0])
            seg_off += qsize(ids.size, b, F)
            prm += ids.size
        np.testing.assert_array_equal(np.sort(np.concatenate(seen)), np.sort(recv_idx[p]))
