"""Exact host solver for AdaQP's bi-objective bit-width assignment -- no PuLP / Gurobi / CBC.

The reference builds, per layer, a MIP (AdaQP/assigner/assigner.py:312-431): binaries
x[b, g] choosing one bit-width b in (2, 4, 8) for every message group g of every channel
c = 'src_dst'; per communication round r an auxiliary Z_r >= alpha_c * MB_c(x) + beta_c for
the W channels (rank -> (rank + r) % W) of that round (:366-377); objective

    lambda * (Var - Var_utopia) / (Var_nadir - Var_utopia)
  + (1 - lambda) * (sum_r Z_r - T_utopia) / (T_nadir - T_utopia)                       (:411)

with Var = sum_c sum_g x[b, g] * var_matrix_c[b, g] and MB_c = sum_g x[b, g] * comm_c[b, g].

Structure exploited (an exact reformulation, not a heuristic):
 1. inside a channel every group costs the same bytes per bit-width (comm_c[b, :] is
    constant, :204-209) and var_matrix_c[b, g] = cost_b * v_g with cost_b decreasing in b;
    by an exchange argument an optimal solution gives the widest bits to the groups with
    the largest v_g, so a channel's choices reduce to (n8, n4): O(G^2) points;
 2. every channel belongs to exactly one round, so the problem separates per round into
    min_Z  a * sum_{c in r} Vmin_c(Z) + b * Z, where Vmin_c(Z) is the channel's minimal
    variance subject to alpha_c * MB_c + beta_c <= Z: a step function of Z whose
    breakpoints are the channels' achievable times.  Scanning the breakpoints gives the
    global optimum.
`brute_force` enumerates all assignments of tiny instances and is used by the tests to
confirm optimality (no MIP solver is installed to compare with).

Schedules.  'ring' is the reference's model of the gloo transport: W-1 rounds, round r
carries the channels rank -> (rank + r) % W and costs the slowest of them, the epoch pays
sum_r Z_r.  'concurrent' models the P2P transport of this repo: ONE send launch per rank
writes all of its peers at once, so a rank's exchange costs alpha_s * (MB of ALL its
channels) + beta_s (alpha/beta fitted on the real send + receive kernel pair,
assigner/profile.py) and the layer pays max over ranks: a single "round" whose units are
the sender ranks.  Pooling a sender's channels keeps the structure above, because groups
of one layer cost the same bytes per bit-width in every channel.
"""
from __future__ import annotations

import itertools
from typing import Dict, List

import numpy as np

BITS = (2, 4, 8)


def _rounds(keys, world_size: int, schedule: str) -> List[List[str]]:
    """Groups of units that run concurrently; the layer pays the sum over groups of the slowest unit."""
    if schedule == "ring":
        out = []
        for r in range(1, world_size):
            chans = [f"{rank}_{(rank + r) % world_size}" for rank in range(world_size)]
            out.append([c for c in chans if c in keys])
        return [g for g in out if g]
    if schedule == "concurrent":
        return [sorted(keys, key=lambda k: int(k))]
    raise ValueError(f"unknown schedule {schedule!r}")


def _pool_by_sender(var_matrix, comm_matrix, cost_model):
    """'concurrent' schedule: one unit per sender rank = its channels side by side."""
    by: Dict[str, List[str]] = {}
    for c in var_matrix:
        by.setdefault(c.split("_")[0], []).append(c)
    v = {s: np.concatenate([np.asarray(var_matrix[c], np.float64) for c in cs], axis=1) for s, cs in by.items()}
    m = {s: np.concatenate([np.asarray(comm_matrix[c], np.float64) for c in cs], axis=1) for s, cs in by.items()}
    k = {s: np.asarray(cost_model[cs[0]], np.float64) for s, cs in by.items()}
    return v, m, k, by


def _scales(var_matrix, comm_matrix, cost_model, world_size: int, schedule: str = "ring"):
    """nadir / utopia of both objectives (assigner.py:344-364, 'nadir_utopia' mode)."""
    var_nadir = sum(float(np.sum(v[0])) for v in var_matrix.values())
    var_utopia = sum(float(np.sum(v[-1])) for v in var_matrix.values())
    t_nadir = t_utopia = 0.0
    for group in _rounds(set(var_matrix), world_size, schedule):
        hi, lo = float("-inf"), float("inf")
        for key in group:
            a, b = cost_model[key][0], cost_model[key][1]
            hi = max(hi, a * float(np.sum(comm_matrix[key][-1])) + b)
            lo = min(lo, a * float(np.sum(comm_matrix[key][0])) + b)
        t_nadir += hi
        t_utopia += lo
    return (var_nadir, var_utopia), (t_nadir, t_utopia)


def _channel_frontier(var_c: np.ndarray, comm_c: np.ndarray, alpha: float, beta: float):
    """All (n8, n4) choices of one channel: returns times ascending, the running minimum of
    the variance over those times, and the (n8, n4) achieving it."""
    G = var_c.shape[1]
    order = np.argsort(-var_c[0], kind="stable")          # groups by decreasing variance weight
    v = var_c[:, order]
    P = [np.concatenate([[0.0], np.cumsum(v[i])]) for i in range(3)]   # prefix sums per bit row
    n8, n4 = np.meshgrid(np.arange(G + 1), np.arange(G + 1), indexing="ij")
    ok = (n8 + n4) <= G
    n8, n4 = n8[ok], n4[ok]
    var = P[2][n8] + (P[1][n8 + n4] - P[1][n8]) + (P[0][G] - P[0][n8 + n4])
    mb2, mb4, mb8 = float(comm_c[0, 0]), float(comm_c[1, 0]), float(comm_c[2, 0])
    mb = mb8 * n8 + mb4 * n4 + mb2 * (G - n8 - n4)
    t = alpha * mb + beta
    idx = np.lexsort((var, t))
    t, var, n8, n4 = t[idx], var[idx], n8[idx], n4[idx]
    best = np.minimum.accumulate(var)
    arg = np.maximum.accumulate(np.where(var <= best, np.arange(var.size), 0))
    # arg[i] = index of the minimal-variance point among the first i+1 (ties -> latest = more bits)
    return t, best, n8[arg], n4[arg], order


def solve_layer(var_matrix: Dict[str, np.ndarray], comm_matrix: Dict[str, np.ndarray],
                cost_model: Dict[str, np.ndarray], coe_lambda: float, world_size: int, schedule: str = "ring"):
    """Optimal group assignment of one layer: {channel: int32[G] bits}, objective value."""
    if schedule == "concurrent":
        v, m, k, by = _pool_by_sender(var_matrix, comm_matrix, cost_model)
        pooled, objective = _solve_units(v, m, k, coe_lambda, world_size, "concurrent")
        result = {}
        for s, cs in by.items():
            off = 0
            for c in cs:
                G = np.asarray(var_matrix[c]).shape[1]
                result[c] = pooled[s][off:off + G].copy()
                off += G
        return result, objective
    return _solve_units(var_matrix, comm_matrix, cost_model, coe_lambda, world_size, schedule)


def _solve_units(var_matrix, comm_matrix, cost_model, coe_lambda: float, world_size: int, schedule: str):
    (vn, vu), (tn, tu) = _scales(var_matrix, comm_matrix, cost_model, world_size, schedule)
    a = coe_lambda / (vn - vu) if vn > vu else 0.0
    b = (1.0 - coe_lambda) / (tn - tu) if tn > tu else 0.0
    result: Dict[str, np.ndarray] = {}
    objective = -a * vu - b * tu
    for chans in _rounds(set(var_matrix), world_size, schedule):
        fr = {c: _channel_frontier(np.asarray(var_matrix[c], np.float64), np.asarray(comm_matrix[c], np.float64),
                                   float(cost_model[c][0]), float(cost_model[c][1])) for c in chans}
        z_min = max(f[0][0] for f in fr.values())
        cand = np.unique(np.concatenate([f[0] for f in fr.values()]))
        cand = cand[cand >= z_min - 1e-15]
        best_obj, best_z = None, None
        for z in cand:
            tot = 0.0
            for c in chans:
                t, vbest = fr[c][0], fr[c][1]
                k = np.searchsorted(t, z + 1e-12, side="right") - 1
                tot += vbest[k]
            obj = a * tot + b * z
            if best_obj is None or obj < best_obj - 1e-15:
                best_obj, best_z = obj, z
        objective += best_obj
        for c in chans:
            t, _, n8s, n4s, order = fr[c]
            k = np.searchsorted(t, best_z + 1e-12, side="right") - 1
            n8, n4 = int(n8s[k]), int(n4s[k])
            G = order.size
            bits_sorted = np.full(G, 2, np.int32)
            bits_sorted[:n8] = 8
            bits_sorted[n8:n8 + n4] = 4
            out = np.empty(G, np.int32)
            out[order] = bits_sorted
            result[c] = out
    return result, objective


def objective_value(assign: Dict[str, np.ndarray], var_matrix, comm_matrix, cost_model, coe_lambda, world_size,
                    schedule: str = "ring"):
    if schedule == "concurrent":
        v, m, k, by = _pool_by_sender(var_matrix, comm_matrix, cost_model)
        pooled = {s: np.concatenate([np.asarray(assign[c]) for c in cs]) for s, cs in by.items()}
        return _objective_units(pooled, v, m, k, coe_lambda, world_size, "concurrent")
    return _objective_units(assign, var_matrix, comm_matrix, cost_model, coe_lambda, world_size, schedule)


def _objective_units(assign, var_matrix, comm_matrix, cost_model, coe_lambda, world_size, schedule):
    (vn, vu), (tn, tu) = _scales(var_matrix, comm_matrix, cost_model, world_size, schedule)
    row = {2: 0, 4: 1, 8: 2}
    var = 0.0
    times = {}
    for c, bits in assign.items():
        rows = np.array([row[int(x)] for x in bits])
        g = np.arange(len(bits))
        var += float(np.sum(np.asarray(var_matrix[c])[rows, g]))
        mb = float(np.sum(np.asarray(comm_matrix[c])[rows, g]))
        times[c] = cost_model[c][0] * mb + cost_model[c][1]
    zsum = 0.0
    for group in _rounds(set(times), world_size, schedule):
        zsum += max(times[c] for c in group)
    a = coe_lambda / (vn - vu) if vn > vu else 0.0
    b = (1.0 - coe_lambda) / (tn - tu) if tn > tu else 0.0
    return a * (var - vu) + b * (zsum - tu)


def brute_force(var_matrix, comm_matrix, cost_model, coe_lambda, world_size, schedule: str = "ring"):
    """Enumerate every assignment (tests only; 3^(total groups) candidates)."""
    chans = list(var_matrix)
    sizes = [np.asarray(var_matrix[c]).shape[1] for c in chans]
    best, best_assign = None, None
    for combo in itertools.product(BITS, repeat=sum(sizes)):
        assign, k = {}, 0
        for c, n in zip(chans, sizes):
            assign[c] = np.array(combo[k:k + n], np.int32)
            k += n
        val = objective_value(assign, var_matrix, comm_matrix, cost_model, coe_lambda, world_size, schedule)
        if best is None or val < best - 1e-15:
            best, best_assign = val, assign
    return best_assign, best
