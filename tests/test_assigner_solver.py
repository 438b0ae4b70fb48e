"""The Gurobi/PuLP-free solver returns optimal solutions of the reference's MIP
(assigner.py:312-431): checked against exhaustive enumeration on small instances."""
import numpy as np
import pytest

from adaqp_b200.assigner import solver

COST = np.array([1 / (2 ** b - 1) ** 2 for b in (2, 4, 8)])


def instance(W, groups, rng, dim=64, gs=10):
    var, com, model = {}, {}, {}
    for s in range(W):
        for d in range(W):
            if s == d:
                continue
            key = f"{s}_{d}"
            G = groups[(s * W + d) % len(groups)]
            v = np.sort(rng.gamma(1.0, 1.0, G))[::-1]
            var[key] = COST[:, None] * v[None, :]
            mb = np.array([2, 4, 8], float)[:, None] * dim * gs / 8 / 2 ** 20
            com[key] = np.repeat(mb, G, axis=1)
            model[key] = np.array([rng.uniform(50, 400), rng.uniform(0.001, 0.05)])
    return var, com, model


@pytest.mark.parametrize("W,groups,lam", [(2, [4, 3], 0.5), (2, [1, 5], 0.2), (3, [2, 1, 2, 1, 1, 2], 0.5),
                                          (3, [1], 0.8), (2, [3, 3], 0.0), (2, [2, 4], 1.0)])
def test_matches_bruteforce(W, groups, lam):
    rng = np.random.default_rng(W * 100 + len(groups) + int(lam * 10))
    var, com, model = instance(W, groups, rng)
    got, obj = solver.solve_layer(var, com, model, lam, W)
    want_assign, want = solver.brute_force(var, com, model, lam, W)
    val = solver.objective_value(got, var, com, model, lam, W)
    assert abs(val - obj) < 1e-9
    assert val <= want + 1e-9, (val, want)
    for key, bits in got.items():
        assert bits.dtype == np.int32 and set(bits.tolist()) <= {2, 4, 8}
        assert bits.shape == (var[key].shape[1],)


def test_extremes():
    rng = np.random.default_rng(1)
    var, com, model = instance(2, [6, 6], rng)
    all8, _ = solver.solve_layer(var, com, model, 1.0, 2)      # variance only -> widest bits
    assert all(np.all(b == 8) for b in all8.values())
    all2, _ = solver.solve_layer(var, com, model, 0.0, 2)      # time only -> narrowest bits on the critical channel
    z = max(model[k][0] * com[k][0].sum() + model[k][1] for k in var)
    for k, b in all2.items():
        t = model[k][0] * sum(com[k][{2: 0, 4: 1, 8: 2}[int(x)], i] for i, x in enumerate(b)) + model[k][1]
        assert t <= z + 1e-12


@pytest.mark.parametrize("W,groups,lam", [(2, [2, 3], 0.5), (3, [1, 1, 2, 1, 1, 1], 0.5), (3, [1], 0.3), (2, [3, 2], 0.9)])
def test_concurrent_schedule_matches_bruteforce(W, groups, lam):
    """P2P transport model: one send launch per rank writes all peers, the layer pays the slowest RANK
    (alpha_rank * MB of all its channels + beta_rank) -- solver vs exhaustive enumeration."""
    rng = np.random.default_rng(7 * W + len(groups) + int(lam * 10))
    var, com, model = instance(W, groups, rng)
    for s in range(W):                       # all channels of a sender share the sender's (alpha, beta)
        ab = np.array([rng.uniform(50, 400), rng.uniform(0.001, 0.05)])
        for d in range(W):
            if s != d:
                model[f"{s}_{d}"] = ab
    got, obj = solver.solve_layer(var, com, model, lam, W, schedule="concurrent")
    _, want = solver.brute_force(var, com, model, lam, W, schedule="concurrent")
    val = solver.objective_value(got, var, com, model, lam, W, schedule="concurrent")
    assert abs(val - obj) < 1e-9
    assert val <= want + 1e-9, (val, want)
    assert set(got) == set(var)
    for key, bits in got.items():
        assert bits.shape == (var[key].shape[1],) and set(bits.tolist()) <= {2, 4, 8}


def test_concurrent_differs_from_ring_when_it_should():
    """Three ranks, rank 0 sends far more than the others: under the ring model every round is bounded by
    a rank-0 channel; under the concurrent model only rank 0's total matters, so ranks 1 and 2 keep 8 bits."""
    rng = np.random.default_rng(3)
    var, com, model = {}, {}, {}
    for s in range(3):
        for d in range(3):
            if s == d:
                continue
            G = 6 if s == 0 else 1
            v = np.sort(rng.gamma(1.0, 1.0, G))[::-1]
            var[f"{s}_{d}"] = COST[:, None] * v[None, :]
            com[f"{s}_{d}"] = np.repeat(np.array([2, 4, 8], float)[:, None] * 64 * 10 / 8 / 2 ** 20, G, axis=1)
            model[f"{s}_{d}"] = np.array([200.0, 0.001])
    got, _ = solver.solve_layer(var, com, model, 0.5, 3, schedule="concurrent")
    assert all(np.all(got[f"{s}_{d}"] == 8) for s in (1, 2) for d in range(3) if d != s)
    assert any(np.any(got[f"0_{d}"] < 8) for d in (1, 2))
