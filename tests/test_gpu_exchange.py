"""GPU parity of the fused exchange kernels: W ranks simulated on one device (their slabs
address each other directly, the same stores a peer GPU would receive over NVLink) against
the CPU oracle.  Packed bytes + bf16 params + generator advance: bit-exact.  Dequantised
halo rows: bit-exact (same IEEE ops).  Also the single-codec <-> fused equivalence."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from adaqp_b200 import build
    build.build()
    from adaqp_b200.helper import DistGNNType
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import SynthSpec
    return dict(DistGNNType=DistGNNType, prepare=prepare_all_in_process, SynthSpec=SynthSpec)


def make_world(env, W, n, deg, F_list, seed=0, chi=0.3):
    from adaqp_b200.communicator.p2p import PeerExchange, wire_in_process
    spec = env["SynthSpec"](name="t", num_nodes=n, num_edges=n * deg, num_parts=W, num_feats=F_list[0],
                            num_classes=5, cross_fraction=chi, community_size=64, seed=seed)
    lays = env["prepare"](spec)
    dev = torch.device("cuda:0")
    exs = [PeerExchange(L.rank, W, dev, F_list, L.send_idx, {p: torch.from_numpy(v) for p, v in L.recv_idx.items()},
                        torch.from_numpy(L.total_send_idx), L.n_halo, timeout_ns=5_000_000_000) for L in lays]
    wire_in_process(exs)
    return spec, lays, exs, dev


def oracle_views(lays):
    send_idx = [L.send_idx for L in lays]
    recv_idx = [L.recv_idx for L in lays]
    nrem = [L.n_halo for L in lays]
    return send_idx, recv_idx, nrem


@pytest.mark.parametrize("W,F", [(2, 100), (4, 256), (3, 602), (4, 13)])
def test_fp32_exchange(env, W, F):
    spec, lays, exs, dev = make_world(env, W, 1200, 10, [F, 32, 32])
    rng = np.random.RandomState(W + F)
    xs = [rng.standard_normal((L.n_inner, F)).astype(np.float32) for L in lays]
    xt = [torch.from_numpy(x).to(dev) for x in xs]
    for rep in range(3):       # repeated use of the same key exercises the ack protocol
        for e, x in zip(exs, xt):
            e.post_send_fp("test0", x)
        halos = [e.complete_recv_fp("test0").clone() for e in exs]
        for e in exs:
            e.release_fp("test0")
    torch.cuda.synchronize()
    for e in exs:
        e.check_status()
    send_idx, recv_idx, nrem = oracle_views(lays)
    want = O.exchange_fp([x[L.total_send_idx] for x, L in zip(xs, lays)], send_idx, recv_idx, nrem)
    for h, w in zip(halos, want):
        np.testing.assert_array_equal(h.cpu().numpy(), w)
    # msg_all2all signature: already gathered send_messages
    for e, x, L in zip(exs, xt, lays):
        e.post_send_fp("forward0", x[torch.from_numpy(L.total_send_idx).to(dev)].contiguous(), gathered=True)
    for e, w in zip(exs, want):
        np.testing.assert_array_equal(e.complete_recv_fp("forward0").cpu().numpy(), w)
    for e in exs:
        e.close()


def random_assignment(lays, keys, rng, uniform=None):
    out = []
    for L in lays:
        a = {}
        for k in keys:
            a[k] = {}
            for p, (lo, hi) in L.send_idx.items():
                if uniform:
                    a[k][p] = torch.full((hi - lo,), uniform, dtype=torch.int32)
                else:
                    a[k][p] = torch.from_numpy(np.array([2, 4, 8], np.int32)[rng.randint(0, 3, hi - lo)])
        out.append(a)
    return out


@pytest.mark.parametrize("W,F,uniform", [(2, 100, None), (4, 256, None), (3, 602, None), (4, 256, 4),
                                         (2, 256, 2), (2, 300, 8), (4, 13, None), (2, 200, None)])
def test_quant_exchange_bit_exact(env, W, F, uniform):
    from adaqp_b200.communicator.p2p import update_quant_in_process
    spec, lays, exs, dev = make_world(env, W, 1500, 10, [F, 64, 64], seed=W)
    rng = np.random.RandomState(10 * W + F)
    key = "forward0"
    assign = random_assignment(lays, [key], rng, uniform)
    update_quant_in_process(exs, assign)
    xs = [np.maximum(rng.standard_normal((L.n_inner, F)), 0).astype(np.float32) for L in lays]
    for x in xs:
        x[::17] = 0.0            # constant rows: scale = inf
    xt = [torch.from_numpy(x).to(dev) for x in xs]
    seeds = [1000 + r for r in range(W)]
    offs = [8 * r for r in range(W)]
    traces = [torch.zeros(L.total_send_idx.size, device=dev) for L in lays]
    for e, x, s, o, t in zip(exs, xt, seeds, offs, traces):
        e.post_send_quant(key, x, s, o, trace=t)
    halos = [e.complete_recv_quant(key).clone() for e in exs]
    torch.cuda.synchronize()
    for e in exs:
        e.check_status()
    send_idx, recv_idx, nrem = oracle_views(lays)
    o_assign = [{p: a[key][p].numpy() for p in a[key]} for a in assign]
    sends = [x[L.total_send_idx] for x, L in zip(xs, lays)]
    want, wire, new_offs = O.exchange_quant(sends, send_idx, recv_idx, nrem, o_assign, seeds, offs, return_wire=True)
    for r, (h, w) in enumerate(zip(halos, want)):
        np.testing.assert_array_equal(h.cpu().numpy().view(np.uint32), w.view(np.uint32))
    for r, e in enumerate(exs):
        assert e.quant_plans[key].philox_increment == new_offs[r] - offs[r]
        for p in e.recv_peers:       # what peer p wrote into my slab == reference wire format
            q, prm = e.recv_region(key, p)
            wq, wprm, valid = wire[p][r]
            got = q.cpu().numpy().view(np.uint8)
            assert got.size == wq.size
            np.testing.assert_array_equal(got[valid], wq[valid])
            np.testing.assert_array_equal(prm.view(torch.int16).cpu().numpy().view(np.uint16), wprm)
        # tracing side product: (F / 6) * (max - min)^2 per send row
        s = sends[r]
        tr = (np.float32(F / 6.0) * (s.max(1) - s.min(1)) ** 2).astype(np.float32)
        np.testing.assert_allclose(traces[r].cpu().numpy(), tr, rtol=1e-6)
    # second exchange of the same key (ack protocol) with gathered inputs and fresh offsets
    for e, x, L, s in zip(exs, xt, lays, seeds):
        e.post_send_quant(key, x[torch.from_numpy(L.total_send_idx).to(dev)].contiguous(), s, 4096, gathered=True)
    halos2 = [e.complete_recv_quant(key).clone() for e in exs]
    want2, _ = O.exchange_quant(sends, send_idx, recv_idx, nrem, o_assign, seeds, [4096] * W)
    for h, w in zip(halos2, want2):
        np.testing.assert_array_equal(h.cpu().numpy().view(np.uint32), w.view(np.uint32))
    for e in exs:
        e.check_status()
        e.close()


def test_fused_equals_single_codec_calls(env):
    """The fused sender == the reference's per-(peer, bit) loop of quant_cuda calls
    (op_util.py:189-209) driven through our quant_cuda drop-in with the torch generator."""
    from adaqp_b200 import quant
    from adaqp_b200.communicator.p2p import update_quant_in_process
    W, F = 3, 256
    spec, lays, exs, dev = make_world(env, W, 900, 8, [F, 64, 64], seed=5)
    rng = np.random.RandomState(1)
    key = "forward0"
    assign = random_assignment(lays, [key], rng)
    update_quant_in_process(exs, assign)
    xt = [torch.from_numpy(rng.standard_normal((L.n_inner, F)).astype(np.float32)).to(dev) for L in lays]
    gen = torch.cuda.default_generators[0]
    r = 1
    e, L = exs[r], lays[r]
    torch.cuda.manual_seed(77)
    seed, off = gen.initial_seed(), gen.get_offset()
    e.post_send_quant(key, xt[r], seed, off)
    torch.cuda.synchronize()
    send = xt[r][torch.from_numpy(L.total_send_idx).to(dev)]
    torch.cuda.manual_seed(77)
    for p, (lo, hi) in L.send_idx.items():
        data = send[lo:hi]
        Q, S, M = [], [], []
        for b in (2, 4, 8):
            ids = torch.nonzero(assign[r][key][p] == b).view(-1).to(dev)
            if len(ids) == 0:
                continue
            sub = data[ids]
            rmin, rmax = torch.min(sub, dim=1)[0], torch.max(sub, dim=1)[0]
            scale = (2 ** b - 1) / (rmax - rmin)
            Q.append(quant.pack_single_precision(sub, rmin, rmax, scale, b, True))
            S.append(scale.to(torch.bfloat16))
            M.append(rmin.to(torch.bfloat16))
        q_ref, prm_ref = torch.concat(Q), torch.stack([torch.concat(S), torch.concat(M)])
        got_q, got_prm = exs[p].recv_region(key, r)
        # compare everything except each segment's unwritten trailing byte
        mask = torch.ones(q_ref.numel(), dtype=torch.bool, device=dev)
        o = 0
        for qq in Q:
            o += qq.numel()
            mask[o - 1] = False
        assert torch.equal(got_q[mask], q_ref[mask])
        assert torch.equal(got_prm.view(torch.int16), prm_ref.view(torch.int16))
    assert gen.get_offset() - off == e.quant_plans[key].philox_increment
    for e in exs:
        e.close()


def test_dead_peer_times_out_and_raises(env):
    """A peer that never sends / never acks: the bounded spins give up after timeout_ns, set the status word, and
    check_status() -- which train_for_one_epoch and val_test poll -- raises instead of letting stale halos through."""
    spec, lays, exs, dev = make_world(env, 2, 600, 8, [32, 32, 32])
    for e in exs:
        e.timeout_ns = 50_000_000                         # 50 ms
    x0 = torch.randn(lays[0].n_inner, 32, device=dev)
    exs[0].post_send_fp("test0", x0)                      # rank 1 never posts its send
    exs[0].complete_recv_fp("test0")
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="flag wait timed out"):
        exs[0].check_status()
    exs[1].check_status()                                 # the silent rank has nothing to report
    exs[0].status.zero_()
    exs[0].post_send_fp("test1", x0)
    exs[0].post_send_fp("test1", x0)                      # rank 1 never consumed the first payload: no ack
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="ack wait timed out"):
        exs[0].check_status()
    for e in exs:
        e.close()
