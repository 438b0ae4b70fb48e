"""Reference data path restated for timing and end-to-end cross-checks -- TEST INFRASTRUCTURE.

`bench.py --impl reference` runs THIS, never the product package: the reference's own flow
for one training epoch (paths relative to /root/reference):

  gather send rows                         AdaQP/model/ops.py:134,164
  per peer x per bit-width quantise loop   AdaQP/model/op_util.py:189-209   (reference quant_cuda
                                           built into oracle/_ref by oracle/build.py)
  D2H into pinned buffers, gloo ring       AdaQP/communicator/comm.py:166-222
  H2D, per peer x per bit dequantise loop  AdaQP/model/op_util.py:211-236
  helper thread + side stream overlap      AdaQP/model/ops.py:119-130,156-193, graphEngine.py:122-132
  Timer = stream synchronise + wall clock  AdaQP/util/timer.py:18-27
  GCN layers / epoch step                  AdaQP/model/distGCN.py:40-85, trainer/runtime_util.py:80-116

Substitutions (DGL is not installable here): graph.update_all(copy_src, sum) -> torch.sparse
CSR SpMM (cuSPARSE) with the same norm multiplies; partitions -> the seeded synthetic
generator shared with the product so both arms see identical inputs.  The gradient
all-reduce is the reference's per-parameter gloo all_reduce on CUDA tensors (:71-77).
"""
from __future__ import annotations

import os
import time
from multiprocessing import Event
from multiprocessing.pool import ThreadPool
from queue import Queue
from typing import Dict

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import build as obuild

BITS_SET = (2, 4, 8)


class RefTimer:
    def __init__(self, device):
        self.device, self.rec = device, {}

    def record(self, name):
        timer = self

        class _Ctx:
            def __enter__(self):
                torch.cuda.current_stream(timer.device).synchronize()
                self.t0 = time.time()

            def __exit__(self, *a):
                torch.cuda.current_stream(timer.device).synchronize()
                timer.rec[name] = timer.rec.get(name, 0.0) + time.time() - self.t0
        return _Ctx()

    def buckets(self):
        out = {"comm": 0.0, "quant": 0.0, "central": 0.0, "marginal": 0.0, "full": 0.0, "exposed": 0.0}
        for k, v in self.rec.items():
            if "communication" in k:
                out["comm"] += v
            elif "quantization" in k:
                out["quant"] += v
            elif "central" in k:
                out["central"] += v
            elif "marginal" in k:
                out["marginal"] += v
            elif "full" in k:
                out["full"] += v
            elif "exposed" in k:
                out["exposed"] += v
        return out

    def clear(self):
        self.rec = {}


class RefState:
    """Per-rank state of the reference's GraphEngine + CommBuffer (restated)."""

    def __init__(self, layout, device, dims, quant: bool, parallel: bool, qc, kind: str = "gcn"):
        L = layout
        self.L, self.dev, self.dims, self.quant, self.parallel, self.qc = L, device, dims, quant, parallel, qc
        self.kind = kind                                          # 'gcn' | 'sage' (mean aggregator)
        self.rank, self.W = L.rank, L.world_size
        ind = torch.from_numpy(L.in_degrees).float().clamp(min=1).to(device)
        outd = torch.from_numpy(L.out_degrees).float().clamp(min=1).to(device)
        self.norm = {"in": ind.pow(-0.5), "out": outd.pow(-0.5), "out_-1": torch.pow(outd, -1)}
        # DGL fn.mean divides by the number of messages = in-degree inside the partition graph
        self.local_indeg = torch.from_numpy(np.diff(L.indptr).astype(np.float32)).to(device)
        n_all = L.n_inner + L.n_halo

        def csr(lo, hi):
            ip = torch.from_numpy((L.indptr[lo:hi + 1] - L.indptr[lo]).astype(np.int64))
            ix = torch.from_numpy(L.indices[L.indptr[lo]:L.indptr[hi]].astype(np.int64))
            return torch.sparse_csr_tensor(ip, ix, torch.ones(ix.numel()), size=(hi - lo, n_all)).to(device)

        self.full = csr(0, L.n_inner)
        if parallel:
            self.central, self.marginal = csr(0, L.n_central), csr(L.n_central, L.n_inner)
        self.total_send_idx = torch.from_numpy(L.total_send_idx).to(device)
        self.send_idx = L.send_idx
        self.recv_idx = {p: torch.from_numpy(v).to(device) for p, v in L.recv_idx.items()}
        self.timer = RefTimer(device)
        self.stream = torch.cuda.Stream(device=device)
        self.quant_ev, self.comp_ev = torch.cuda.Event(), torch.cuda.Event()
        self.quant_cpu, self.comp_cpu = Event(), Event()
        self.pool = ThreadPool(processes=1)
        # fp32 ("test") buffers, buffer.py:154-174
        self.fp_send = [{p: torch.zeros((hi - lo, d)).pin_memory() for p, (lo, hi) in L.send_idx.items()} for d in dims]
        self.fp_recv_cpu = [{p: torch.zeros((len(v), d)).pin_memory() for p, v in L.recv_idx.items()} for d in dims]
        self.fp_recv_gpu = [{p: torch.zeros((len(v), d), device=device) for p, v in L.recv_idx.items()} for d in dims]
        self.q = {}

    @staticmethod
    def qsize(n, b, F):
        wpt = 8 // b
        nr = n + (wpt - n % wpt) % wpt
        return (b * nr * F + 8) // 8

    def update_quant(self, assignment: Dict[str, Dict[int, torch.Tensor]]):
        """buffer.py:176-248 incl. the all_gather_object of idx lists and sizes."""
        send_ids, send_sizes = {}, {}
        for key, per in assignment.items():
            Fd = self.dims[int(key[-1])]
            send_ids[key], send_sizes[key] = {}, {}
            for p, cfg in per.items():
                send_ids[key][p], send_sizes[key][p] = {}, {}
                for b in BITS_SET:
                    ids = torch.nonzero(cfg == b).view(-1)
                    if len(ids):
                        send_ids[key][p][b] = ids
                        send_sizes[key][p][b] = (self.qsize(len(ids), b, Fd), len(ids))
        gathered = [None] * self.W
        dist.all_gather_object(gathered, [send_ids, send_sizes])
        self.q = {}
        for key in assignment:
            ent = {"send_ids": {p: {b: i.to(self.dev) for b, i in d.items()} for p, d in send_ids[key].items()},
                   "send_cpu": {}, "recv_cpu": {}, "recv_gpu": {}, "recv_ids": {}, "recv_sizes": {}}
            for p, sizes in send_sizes[key].items():
                qt, fp = sum(s[0] for s in sizes.values()), sum(s[1] for s in sizes.values())
                ent["send_cpu"][p] = (torch.zeros(qt, dtype=torch.int8).pin_memory(), torch.zeros((2, fp), dtype=torch.bfloat16).pin_memory())
            for i in range(self.W):
                if i != self.rank and self.rank in gathered[i][0][key]:
                    ent["recv_ids"][i] = {b: v.to(self.dev) for b, v in gathered[i][0][key][self.rank].items()}
                    ent["recv_sizes"][i] = gathered[i][1][key][self.rank]
                    sizes = ent["recv_sizes"][i]
                    qt, fp = sum(s[0] for s in sizes.values()), sum(s[1] for s in sizes.values())
                    ent["recv_cpu"][i] = (torch.zeros(qt, dtype=torch.int8).pin_memory(), torch.zeros((2, fp), dtype=torch.bfloat16).pin_memory())
                    ent["recv_gpu"][i] = (torch.zeros(qt, dtype=torch.int8, device=self.dev), torch.zeros((2, fp), dtype=torch.bfloat16, device=self.dev))
            self.q[key] = ent

    # ---- exchanges (comm.py:166-222) -------------------------------------------------------
    def fp_exchange(self, li, send_messages):
        rank, W = self.rank, self.W
        sends, recvs = [], Queue()
        for i in range(1, W):
            left, right = (rank - i + W) % W, (rank + i) % W
            lo, hi = self.send_idx[right]
            self.fp_send[li][right].copy_(send_messages[lo:hi])
            sends.append(dist.isend(self.fp_send[li][right], right, tag=0))
            recvs.put((dist.irecv(self.fp_recv_cpu[li][left], left, tag=0), left))
        while not recvs.empty():
            r, left = recvs.get()
            r.wait()
            self.fp_recv_gpu[li][left].copy_(self.fp_recv_cpu[li][left], non_blocking=True)
        for r in sends:
            r.wait()

    def qt_exchange(self, ent):
        rank, W = self.rank, self.W
        sends, recvs = [], Queue()
        for i in range(1, W):
            left, right = (rank - i + W) % W, (rank + i) % W
            qd, qp = ent["send_cpu"][right]
            sends += [dist.isend(qd, right, tag=0), dist.isend(qp, right, tag=1)]
            recvs.put((dist.irecv(ent["recv_cpu"][left][0], left, tag=0), dist.irecv(ent["recv_cpu"][left][1], left, tag=1), left))
        while not recvs.empty():
            r0, r1, left = recvs.get()
            r0.wait()
            r1.wait()
            ent["recv_gpu"][left][0].copy_(ent["recv_cpu"][left][0], non_blocking=True)
            ent["recv_gpu"][left][1].copy_(ent["recv_cpu"][left][1], non_blocking=True)
        for r in sends:
            r.wait()

    # ---- op_util.py:137-236 -------------------------------------------------------------------------
    def all2all(self, send_messages, name, is_train):
        side = self.parallel
        ctx = torch.cuda.stream(self.stream) if side else _Null()
        with ctx:
            Fd = send_messages.shape[1]
            n_rem = self.L.n_halo
            if not (self.quant and is_train):
                li = int(name[-1])
                with self.timer.record(f"{name}_communication"):
                    self.fp_exchange(li, send_messages)
                remote = torch.zeros(n_rem, Fd, device=self.dev)
                for p, idx in self.recv_idx.items():
                    remote[idx] = self.fp_recv_gpu[li][p]
                return remote
            ent = self.q[name]
            with self.timer.record(f"{name}_quantization"):
                for p, (lo, hi) in self.send_idx.items():
                    data = send_messages[lo:hi]
                    Q, S, M = [], [], []
                    for b, ids in ent["send_ids"][p].items():
                        sub = data[ids]
                        rmin, rmax = torch.min(sub, dim=1)[0], torch.max(sub, dim=1)[0]
                        scale = (2 ** b - 1) / (rmax - rmin)
                        Q.append(self.qc.pack_single_precision(sub, rmin, rmax, scale.to(sub.dtype), b, True))
                        S.append(scale.to(torch.bfloat16))
                        M.append(rmin.to(torch.bfloat16))
                    ent["send_cpu"][p][0].copy_(torch.concat(Q), non_blocking=True)
                    ent["send_cpu"][p][1].copy_(torch.stack([torch.concat(S), torch.concat(M)], dim=0), non_blocking=True)
                if side:
                    self.quant_ev.record(torch.cuda.current_stream())
                    self.quant_cpu.set()
            with self.timer.record(f"{name}_communication"):
                self.qt_exchange(ent)
            with self.timer.record(f"{name}_de-quantization"):
                if side:
                    self.comp_cpu.wait()
                    torch.cuda.current_stream().wait_event(self.comp_ev)
                    self.quant_cpu.clear()
                remote = torch.zeros(n_rem, Fd, device=self.dev)
                for p, ids in self.recv_idx.items():
                    qd, qp = ent["recv_gpu"][p]
                    sub = remote[ids]
                    qo = fo = 0
                    for b, (qs, n) in ent["recv_sizes"][p].items():
                        orig = ent["recv_ids"][p][b]
                        sc, mn = qp[0, fo:fo + n].to(torch.float32), qp[1, fo:fo + n].to(torch.float32)
                        sub[orig] = self.qc.unpack_single_precision(qd[qo:qo + qs].contiguous(), b, sc, mn, len(orig), Fd).contiguous()
                        qo += qs
                        fo += n
                    remote[ids] = sub
            return remote

    # ---- ops.py:17-32 with torch.sparse in place of DGL update_all ---------------------------------------
    def gcn_agg(self, csr, feats, lo, hi, backward):
        n1, n2 = (self.norm["in"], self.norm["out"]) if backward else (self.norm["out"], self.norm["in"])
        h = feats * n1[:feats.shape[0]].view(-1, 1)
        if h.shape[0] < csr.shape[1]:
            h = torch.cat([h, h.new_zeros(csr.shape[1] - h.shape[0], h.shape[1])], 0)
        return torch.sparse.mm(csr, h) * n2[lo:hi].view(-1, 1)

    # ---- ops.py:34-67, aggregator_type='mean' -------------------------------------------------------------
    def sage_agg(self, csr, feats, lo, hi, backward):
        h = feats * self.norm["out_-1"][:feats.shape[0]].view(-1, 1) if backward else feats
        if h.shape[0] < csr.shape[1]:
            h = torch.cat([h, h.new_zeros(csr.shape[1] - h.shape[0], h.shape[1])], 0)
        out = torch.sparse.mm(csr, h)
        if not backward:
            out = out / self.local_indeg[lo:hi].clamp(min=1).view(-1, 1)
        return out

    def agg(self, csr, feats, lo, hi, backward):
        return (self.gcn_agg if self.kind == "gcn" else self.sage_agg)(csr, feats, lo, hi, backward)

    def propagate(self, x, layer, is_train, backward):
        name = f"backward{layer}" if backward else f"forward{layer}"
        L = self.L
        if self.W == 1:
            with self.timer.record(f"{name}_full_aggregation"):
                return self.agg(self.full, x, 0, L.n_inner, backward)
        if not self.parallel:                                  # ops.py:132-154
            send = x[self.total_send_idx]
            remote = self.all2all(send, name, is_train)
            full = torch.cat([x, remote], dim=0)
            with self.timer.record(f"{name}_full_aggregation"):
                return self.agg(self.full, full, 0, L.n_inner, backward)
        torch.cuda.current_stream().synchronize()               # ops.py:162
        send = x[self.total_send_idx]
        resp = self.pool.apply_async(self.all2all, args=(send, name, is_train))
        q = self.quant and is_train
        if q:                                                   # central_compute_ctx, ops.py:119-130
            self.quant_cpu.wait()
            torch.cuda.current_stream().wait_event(self.quant_ev)
        with self.timer.record(f"{name}_central_aggregation"):
            cen = self.agg(self.central, x, 0, L.n_central, backward)
        if q:
            self.comp_ev.record(torch.cuda.current_stream())
            self.comp_cpu.set()                                 # never cleared, as in the reference
        t0 = time.time()
        remote = resp.get()                                     # exposed comm wait, ops.py:177
        torch.cuda.current_stream().wait_stream(self.stream)
        self.timer.rec[f"{name}_exposed"] = self.timer.rec.get(f"{name}_exposed", 0.0) + time.time() - t0
        full = torch.cat([x, remote], dim=0)
        with self.timer.record(f"{name}_marginal_aggregation"):
            mar = self.agg(self.marginal, full, L.n_central, L.n_inner, backward)
        return torch.cat([cen, mar], dim=0)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _RefAgg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, state, layer, is_train):
        ctx.state, ctx.layer = state, layer
        return state.propagate(x, layer, is_train, False)

    @staticmethod
    def backward(ctx, g):
        return ctx.state.propagate(g.contiguous(), ctx.layer, True, True), None, None, None


class RefGCN(nn.Module):
    """distGCN.py:52-85 (aggregate, matmul, bias; dropout -> LayerNorm -> ReLU between layers)."""

    def __init__(self, dims, classes, drop):
        super().__init__()
        sizes = dims + [classes]
        self.w = nn.ParameterList([nn.Parameter(nn.init.xavier_uniform_(torch.empty(sizes[i], sizes[i + 1]))) for i in range(len(dims))])
        self.b = nn.ParameterList([nn.Parameter(torch.zeros(sizes[i + 1])) for i in range(len(dims))])
        self.norms = nn.ModuleList([nn.LayerNorm(sizes[i + 1]) for i in range(len(dims) - 1)])
        self.drop = drop

    def forward(self, state, x):
        n = len(self.w)
        for i in range(n):
            x = torch.matmul(_RefAgg.apply(x, state, i, self.training), self.w[i]) + self.b[i]
            if i < n - 1:
                x = F.relu(self.norms[i](F.dropout(x, p=self.drop, training=self.training)), inplace=True)
        return x


class RefSAGE(nn.Module):
    """distSAGE.py:14-97, aggregator 'mean': fc_self(h) + fc_neigh(mean of in-neighbours) + bias."""

    def __init__(self, dims, classes, drop):
        super().__init__()
        sizes = dims + [classes]
        gain = nn.init.calculate_gain("relu")
        n = len(dims)
        self.fc_self = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1], bias=False) for i in range(n)])
        self.fc_neigh = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1], bias=False) for i in range(n)])
        for m in list(self.fc_self) + list(self.fc_neigh):
            nn.init.xavier_uniform_(m.weight, gain=gain)
        self.b = nn.ParameterList([nn.Parameter(torch.zeros(sizes[i + 1])) for i in range(n)])
        self.norms = nn.ModuleList([nn.LayerNorm(sizes[i + 1]) for i in range(n - 1)])
        self.drop = drop

    def forward(self, state, x):
        n = len(self.b)
        for i in range(n):
            h = _RefAgg.apply(x, state, i, self.training)
            x = self.fc_self[i](x) + self.fc_neigh[i](h) + self.b[i]
            if i < n - 1:
                x = F.relu(self.norms[i](F.dropout(x, p=self.drop, training=self.training)))
        return x


def make_model(kind: str, dims, classes, drop):
    return RefGCN(dims, classes, drop) if kind == "gcn" else RefSAGE(dims, classes, drop)


def load_product_state(ref: nn.Module, state: Dict[str, torch.Tensor]):
    """Copy a product DistGCN / DistSAGE state_dict (reference parameter names, distGCN.py:52-75,
    distSAGE.py:62-80) into the restated model, for the same-weights activation cross-check."""
    with torch.no_grad():
        if isinstance(ref, RefGCN):
            for i in range(len(ref.w)):
                ref.w[i].copy_(state[f"convs.{i}.weight"])
                ref.b[i].copy_(state[f"convs.{i}.bias"])
        else:
            for i in range(len(ref.b)):
                ref.fc_self[i].weight.copy_(state[f"sages.{i}.fc_self.weight"])
                ref.fc_neigh[i].weight.copy_(state[f"sages.{i}.fc_neigh.weight"])
                ref.b[i].copy_(state[f"sages.{i}.bias"])
        for i in range(len(ref.norms)):
            ref.norms[i].weight.copy_(state[f"norms.{i}.weight"])
            ref.norms[i].bias.copy_(state[f"norms.{i}.bias"])


def bench(args, rank, world):
    """K training epochs of the reference flow; same JSON contract as the product arm."""
    import yaml
    if not torch.cuda.is_available():
        return {"impl": "reference", "unavailable": "no CUDA device for the reference's GPU-side kernels"}
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="env://")
    quant = args.mode in ("AdaQP", "AdaQP-q")
    parallel = args.mode in ("AdaQP", "AdaQP-p") and world > 1
    qc = None
    if quant and world > 1:
        if not obuild.ref_available():
            return {"impl": "reference", "unavailable": "oracle/_ref/quant_cuda.so (reference kernels) not built"}
        qc = obuild.load_ref()
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(here, "adaqp_b200", "config", f"{args.dataset}.yaml")))
    # the partition generator is shared so that both arms see identical inputs
    from adaqp_b200.helper import DistGNNType
    from adaqp_b200.manager.layout import prepare_rank
    from adaqp_b200.manager.partition_synth import spec_from_config
    spec = spec_from_config(cfg, world, args.scale)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    kind = args.model_name
    L = getattr(args, "layout", None)                      # a layout already prepared in this process (run_configs)
    if L is None:
        L = prepare_rank(spec, rank, DistGNNType.DistGCN if kind == "gcn" else DistGNNType.DistSAGE, gather)
    dims = [cfg["data"]["num_feats"]] + [cfg["model"]["hidden_dim"]] * (cfg["model"]["num_layers"] - 1)
    st = RefState(L, dev, dims, quant, parallel, qc, kind)
    torch.manual_seed(2024)
    torch.cuda.manual_seed(2024)
    if quant and world > 1:
        keys = [f"forward{i}" for i in range(len(dims))] + [f"backward{i}" for i in range(1, len(dims))]
        bits = torch.tensor(BITS_SET, dtype=torch.int32)
        ubits = getattr(args, "assign_bits", None) or cfg["assignment"]["assign_bits"]
        given = getattr(args, "assignment", None)          # e.g. the product arm's adaptive result (tools/run_configs.py)
        if given is not None:
            assign = given
        else:                                              # 'adaptive' without a given result: random {2,4,8} stands in
            assign = {k: {p: (torch.full((hi - lo,), ubits, dtype=torch.int32) if args.assign_scheme == "uniform"
                              else bits[torch.multinomial(torch.full((3,), 1 / 3), hi - lo, replacement=True)])
                          for p, (lo, hi) in L.send_idx.items()} for k in keys}
        st.update_quant(assign)
    model = make_model(kind, dims, cfg["data"]["num_classes"], cfg["model"]["dropout_rate"]).to(dev)
    for v in model.state_dict().values():
        if rank != 0:
            v.zero_()
        dist.all_reduce(v.data)
    feats = torch.from_numpy(L.feat).to(dev)
    labels = torch.from_numpy(L.label).to(dev)
    train_mask = torch.nonzero(torch.from_numpy(L.train_mask)).squeeze(1).to(dev)
    n_train = torch.LongTensor([train_mask.numel()])
    dist.all_reduce(n_train)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["runtime"]["learning_rate"])
    crit = nn.BCEWithLogitsLoss(reduction="sum") if cfg["data"]["is_multilabel"] else nn.CrossEntropyLoss(reduction="sum")

    def epoch():
        model.train()
        logits = model(st, feats)
        loss = crit(logits[train_mask], labels[train_mask]) / n_train.item()
        opt.zero_grad()
        loss.backward()
        for p in model.parameters():
            dist.all_reduce(p.grad.data)                       # runtime_util.py:71-77
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        epoch()
    st.timer.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(args.steps):
        loss = epoch()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.time() - t0], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    b = st.timer.buckets()
    exposed = (b["exposed"] if parallel else b["comm"] + b["quant"]) / args.steps * 1e3
    ex = torch.tensor([exposed], dtype=torch.float64)
    dist.all_reduce(ex, op=dist.ReduceOp.MAX)
    cores = len(os.sched_getaffinity(0))
    val = args.steps / float(dt.item())
    st.pool.close()
    return {"impl": "reference", "metric": "epochs_per_sec", "value": val, "unit": "epochs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * float(dt.item()) / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "exposed_comm_ms": float(ex.item()),
            "breakdown_s_per_epoch": {k: v / args.steps for k, v in b.items()},
            "config": {"workload": f"{args.dataset}-shape {args.model_name} 3x256 full-graph training epoch, {world} partition(s), "
                                   f"mode {args.mode}, bits {args.assign_scheme}{{2,4,8}}", "layer_dims": dims,
                       "nodes": spec.num_nodes, "edges": spec.num_edges},
            "cpu_baseline": {"value": val, "unit": "epochs/s", "cores": cores, "kind": "port",
                             "sample": "reference flow restated (oracle/ref_path.py): gloo ring over pinned host buffers on the box's host "
                                       f"cores ({cores} usable, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')}), reference-built "
                                       "quant_cuda kernels, torch.sparse SpMM in place of DGL; whole epochs, not a sample"},
            "e2e": {"value": val, "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
