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
