"""Generate golden vectors from the REFERENCE's own quant_cuda kernels.

TEST INFRASTRUCTURE.  Run on a B200 box (needs a GPU and oracle/_ref/quant_cuda*.so,
which oracle/build.py compiles from /root/reference in the build container):

    gpurun -- python -m oracle.make_golden gpurun_out/golden

then copy gpurun_out/golden/*.npz to tests/golden/ and commit them together with this
script.  Each file records inputs, the generator state handed to the kernel and the
reference outputs, so the CPU oracle (oracle/quant_oracle.c) can be pinned without a GPU
and the product kernels can be compared against the same bytes.

Host-side steps around the kernels restate AdaQP/model/op_util.py:20-83,189-236 with
torch ops on the GPU (torch.min/max, fp32 scale, bf16 casts) exactly as the reference
performs them; only quant_cuda.{pack,unpack}_single_precision come from oracle/_ref.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import build as obuild

BITS_SET = (2, 4, 8)

SINGLE_CASES = [
    # name, N, F, bits, data kind
    ("n7_f13_b2", 7, 13, 2, "normal"),
    ("n8_f100_b4", 8, 100, 4, "normal"),
    ("n5_f602_b8", 5, 602, 8, "normal"),
    ("n33_f256_b2", 33, 256, 2, "relu"),
    ("n16_f256_b4", 16, 256, 4, "relu"),
    ("n3_f300_b8", 3, 300, 8, "grad"),
    ("n9_f200_b1", 9, 200, 1, "normal"),
    ("n6_f64_b4_edge", 6, 64, 4, "edge"),
    ("n1_f1_b8", 1, 1, 8, "normal"),
    ("n2_f1024_b2", 2, 1024, 2, "normal"),
]

MIXED_CASES = [
    # name, S, F, seed
    ("mixed_s23_f100", 23, 100, 11),
    ("mixed_s40_f256", 40, 256, 12),
    ("mixed_s9_f602", 9, 602, 13),
    ("mixed_s5_f256_only8", 5, 256, 14),
]


def make_data(kind: str, N: int, F: int, rng: np.random.RandomState) -> np.ndarray:
    if kind == "normal":
        x = rng.standard_normal((N, F))
    elif kind == "relu":
        x = np.maximum(rng.standard_normal((N, F)), 0.0)
    elif kind == "grad":
        x = rng.standard_normal((N, F)) * 1e-3
        x[0] = 0.0  # all-zero gradient row: scale = inf
    elif kind == "edge":
        x = rng.standard_normal((N, F))
        x[0] = 3.25          # constant row -> range 0 -> scale inf
        x[1] = 0.0
        x[2, ::2] = 1e30     # huge range
        x[3] = np.where(rng.rand(F) < 0.5, -1.0, 1.0)  # only min / max values
        x[4, 0] = np.float32(65504.0)
    else:
        raise ValueError(kind)
    return x.astype(np.float32)


def gen_state(dev):
    g = torch.cuda.default_generators[dev.index or 0]
    return g.initial_seed(), g.get_offset()


def run_single(qc, name, N, F, bits, kind, dev, seed):
    rng = np.random.RandomState(seed)
    x = make_data(kind, N, F, rng)
    xt = torch.from_numpy(x).to(dev)
    rmin = torch.min(xt, dim=1)[0]
    rmax = torch.max(xt, dim=1)[0]
    scale = ((2 ** bits - 1) / (rmax - rmin)).to(torch.float32)
    torch.cuda.manual_seed(1234 + seed)
    s0, o0 = gen_state(dev)
    packed = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    s1, o1 = gen_state(dev)
    wpt = 8 // bits
    payload = ((N + wpt - 1) // wpt) * F
    deq = qc.unpack_single_precision(packed, bits, scale, rmin, N, F)
    # wire params as the reference casts them (op_util.py:72-74,79-81)
    sc16 = scale.to(torch.bfloat16)
    mn16 = rmin.to(torch.bfloat16)
    deq16 = qc.unpack_single_precision(packed, bits, sc16.to(torch.float32), mn16.to(torch.float32), N, F)
    torch.cuda.synchronize()
    return dict(
        x=x, bits=np.int32(bits), rmin=rmin.cpu().numpy(), rmax=rmax.cpu().numpy(),
        scale=scale.cpu().numpy(), seed=np.uint64(s0), offset=np.uint64(o0),
        offset_after=np.uint64(o1), packed_len=np.int64(packed.numel()),
        payload=packed[:payload].cpu().numpy().view(np.uint8), deq=deq.cpu().numpy(),
        scale_bf16=sc16.view(torch.int16).cpu().numpy().view(np.uint16),
        min_bf16=mn16.view(torch.int16).cpu().numpy().view(np.uint16),
        deq_bf16=deq16.cpu().numpy())


HALF_CASES = [
    # name, N, F, bits, data kind   (fp16 instantiation of the reference kernels)
    ("h_n7_f13_b2", 7, 13, 2, "normal"),
    ("h_n8_f100_b4", 8, 100, 4, "normal"),
    ("h_n5_f602_b8", 5, 602, 8, "normal"),
    ("h_n33_f256_b2", 33, 256, 2, "relu"),
    ("h_n16_f256_b4", 16, 256, 4, "relu"),
    ("h_n3_f300_b8", 3, 300, 8, "grad"),
    ("h_n9_f200_b1", 9, 200, 1, "normal"),
    ("h_n6_f64_b4_edge", 6, 64, 4, "edge_half"),
]


def run_single_half(qc, name, N, F, bits, kind, dev, seed):
    """op_util.py:24-43 with float16 tensors: min / max / scale are computed by torch in half."""
    rng = np.random.RandomState(seed)
    if kind == "edge_half":
        x = rng.standard_normal((N, F)).astype(np.float32)
        x[0] = 3.25                     # constant row -> scale inf -> NaN -> 0
        x[1] = 0.0
        x[2] = 1.0
        x[2, ::2] = 1.001               # range 0.001 < 15 / 65504: the half scale overflows to inf
        x[3] = np.where(rng.rand(F) < 0.5, -1.0, 1.0)
        x[4, 0] = 60000.0
    else:
        x = make_data(kind, N, F, rng)
    xt = torch.from_numpy(x).to(dev).to(torch.float16)
    rmin = torch.min(xt, dim=1)[0]
    rmax = torch.max(xt, dim=1)[0]
    scale = ((2 ** bits - 1) / (rmax - rmin)).to(xt.dtype)
    torch.cuda.manual_seed(4321 + seed)
    s0, o0 = gen_state(dev)
    packed = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    s1, o1 = gen_state(dev)
    wpt = 8 // bits
    payload = ((N + wpt - 1) // wpt) * F
    deq = qc.unpack_single_precision(packed, bits, scale, rmin, N, F)
    torch.cuda.synchronize()
    assert deq.dtype == torch.float16
    u16 = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    return dict(x=u16(xt), bits=np.int32(bits), rmin=u16(rmin), rmax=u16(rmax), scale=u16(scale), seed=np.uint64(s0),
                offset=np.uint64(o0), offset_after=np.uint64(o1), packed_len=np.int64(packed.numel()),
                payload=packed[:payload].cpu().numpy().view(np.uint8), deq=u16(deq))


def run_mixed(qc, name, S, F, seed, dev):
    """One src->dst channel: mixed_msg_quantization / mixed_msg_dequantization."""
    rng = np.random.RandomState(seed)
    x = np.maximum(rng.standard_normal((S, F)), 0).astype(np.float32)
    if "only8" in name:
        assign = np.full(S, 8, np.int32)
    else:
        assign = np.array(BITS_SET, np.int32)[rng.randint(0, 3, size=S)]
    xt = torch.from_numpy(x).to(dev)
    at = torch.from_numpy(assign)
    torch.cuda.manual_seed(99 + seed)
    s0, o0 = gen_state(dev)
    qparts, valid, scs, mns, sizes = [], [], [], [], []
    for b in BITS_SET:
        ids = torch.nonzero(at == b).view(-1)
        if len(ids) == 0:
            continue
        sub = xt[ids.to(dev)]
        rmin, rmax = torch.min(sub, dim=1)[0], torch.max(sub, dim=1)[0]
        scale = (2 ** b - 1) / (rmax - rmin)
        q = qc.pack_single_precision(sub, rmin, rmax, scale.to(sub.dtype), b, True)
        wpt = 8 // b
        payload = ((len(ids) + wpt - 1) // wpt) * F
        v = np.zeros(q.numel(), bool)
        v[:payload] = True
        qparts.append(q)
        valid.append(v)
        scs.append(scale.to(torch.bfloat16))
        mns.append(rmin.to(torch.bfloat16))
        sizes.append((b, q.numel(), len(ids)))
    s1, o1 = gen_state(dev)
    qdata = torch.concat(qparts)
    params = torch.stack([torch.concat(scs), torch.concat(mns)], dim=0)
    # receiver (op_util.py:216-235)
    out = torch.zeros(S, F, device=dev)
    q_off = fp_off = 0
    for b, qs, n in sizes:
        ids = torch.nonzero(at == b).view(-1).to(dev)
        sc = params[0, fp_off:fp_off + n].to(torch.float32)
        mn = params[1, fp_off:fp_off + n].to(torch.float32)
        out[ids] = qc.unpack_single_precision(qdata[q_off:q_off + qs].contiguous(), b, sc, mn, n, F).contiguous()
        q_off += qs
        fp_off += n
    torch.cuda.synchronize()
    qd = qdata.cpu().numpy().view(np.uint8).copy()
    vm = np.concatenate(valid)
    qd[~vm] = 0
    return dict(x=x, assign=assign, seed=np.uint64(s0), offset=np.uint64(o0),
                offset_after=np.uint64(o1), qdata=qd, valid=vm,
                params=params.view(torch.int16).cpu().numpy().view(np.uint16), deq=out.cpu().numpy())


def main(out_dir: str, half_only: bool = False):
    os.makedirs(out_dir, exist_ok=True)
    qc = obuild.load_ref()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    for i, (name, N, F, bits, kind) in enumerate(HALF_CASES):
        rec = run_single_half(qc, name, N, F, bits, kind, dev, seed=300 + i)
        np.savez_compressed(os.path.join(out_dir, f"half_{name}.npz"), **rec)
        print("golden", name, "offset", int(rec["offset"]), "->", int(rec["offset_after"]))
    if half_only:
        with open(os.path.join(out_dir, "PROVENANCE_half.txt"), "w") as f:
            f.write(f"half_*.npz generated by oracle/make_golden.py --half on {torch.cuda.get_device_name(0)} with torch "
                    f"{torch.__version__}; kernels: reference quant_cuda (fp16 instantiation) built from /root/reference\n")
        return
    for i, (name, N, F, bits, kind) in enumerate(SINGLE_CASES):
        rec = run_single(qc, name, N, F, bits, kind, dev, seed=100 + i)
        np.savez_compressed(os.path.join(out_dir, f"single_{name}.npz"), **rec)
        print("golden", name, "offset", int(rec["offset"]), "->", int(rec["offset_after"]))
    for name, S, F, seed in MIXED_CASES:
        rec = run_mixed(qc, name, S, F, seed, dev)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **rec)
        print("golden", name, "offset", int(rec["offset"]), "->", int(rec["offset_after"]))
    with open(os.path.join(out_dir, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by oracle/make_golden.py on {torch.cuda.get_device_name(0)} with torch "
                f"{torch.__version__}; kernels: reference quant_cuda built from /root/reference "
                f"(oracle/build.py build_ref)\n")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0] if args else "gpurun_out/golden", half_only="--half" in sys.argv)
