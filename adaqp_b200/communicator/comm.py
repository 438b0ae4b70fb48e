"""Communicator: process-group control plane + boundary-message data plane.

API of the reference's Communicator (AdaQP/communicator/comm.py:14-248) kept: ctor
signature, class attribute `ctx`, the static collective / p2p wrappers, the two exchange
methods and the buffer wrappers.  Differences, by design:

* torch.distributed(gloo) is the CONTROL plane only (object collectives, barriers,
  metadata, CUDA-IPC handle exchange).
* On CUDA devices the DATA plane is `transport == 'p2p'`: the fused kernels of
  csrc/exchange.cu store into peer-mapped slabs over NVLink/NVSwitch
  (communicator/p2p.py).  `fp_msg_exchange` / `qt_msg_exchange` remain as the reference's
  gloo ring over host buffers (comm.py:166-222) and are what `transport == 'gloo'` uses:
  the CPU plumbing configuration (ADAQP_DEVICE=cpu) and the timed baseline.
"""
from __future__ import annotations

import logging
import os
from queue import Queue
from typing import Any, Dict, List, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from ..helper import MessageType
from .buffer import Basic_Buffer_Type, CommBuffer

logger = logging.getLogger("trainer")


_EXIT_HOOK = False


def _pick_device(local_rank: int) -> torch.device:
    want = os.environ.get("ADAQP_DEVICE", "").lower()
    if want == "cpu":
        return torch.device("cpu")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path is CUDA-only. Set ADAQP_DEVICE=cpu for the "
                           "gloo plumbing mode (fp32 Vanilla only).")
    return torch.device(f"cuda:{local_rank}")


class Communicator(object):
    ctx: "Communicator" = None

    def __init__(self, backend: str = "gloo", init_method: str = "env://"):
        self._init(backend, init_method)
        self.comm_buffer: CommBuffer = None
        Communicator.ctx = self

    def _init(self, backend: str, init_method: str):
        if backend != "gloo":
            raise NotImplementedError("only gloo is supported now")
        if not dist.is_initialized():
            dist.init_process_group(backend, init_method=init_method)
        global _EXIT_HOOK
        if not _EXIT_HOOK:
            import atexit
            atexit.register(Communicator._destroy)
            _EXIT_HOOK = True
        self._backend = backend
        self._init_method = init_method
        self._local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._device = _pick_device(self._local_rank)
        if self._device.type == "cuda":
            torch.cuda.set_device(self._device)
        transport = os.environ.get("ADAQP_TRANSPORT", "").lower()
        self.transport = transport if transport in ("p2p", "gloo") else ("p2p" if self._device.type == "cuda" else "gloo")
        if self.transport == "p2p" and self._device.type != "cuda":
            raise RuntimeError("transport 'p2p' needs a CUDA device")

    def __repr__(self):
        return (f"<Communicator(rank: {self.get_rank()}, backend: {self.backend}, world_size: "
                f"{self.get_world_size()}, local_rank: {self.local_rank}, device: {self.device}, "
                f"transport: {self.transport})>")

    # ---- getters (comm.py:46-60): local_rank / device / init_method / backend are attached below ------
    @staticmethod
    def get_rank():
        return dist.get_rank()

    @staticmethod
    def get_world_size():
        return dist.get_world_size()

    @staticmethod
    def get_backend():
        return dist.get_backend()

    @staticmethod
    def _destroy():
        if dist.is_initialized():
            dist.destroy_process_group()

    def __del__(self):
        # The reference destroys the process group in __del__ (comm.py:230-233).  Here the group outlives
        # any one communicator object (a process may build several Trainers in a row, and object
        # finalisation order is not under our control): it is destroyed once, at interpreter exit.
        pass

    @staticmethod
    def barrier():
        dist.barrier()

    # ---- collectives (control plane) ----------------------------------------------------
    @staticmethod
    def all_reduce_max(tensor: Tensor):
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)

    @staticmethod
    def all_reduce_sum(tensor: Tensor):
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)

    @staticmethod
    def all_gather_any(obj_list: List[Any], obj: Any):
        dist.all_gather_object(obj_list, obj)

    @staticmethod
    def broadcast_any(obj_list: List[Any], src: int = 0):
        dist.broadcast_object_list(obj_list, src)

    @staticmethod
    def scatter_any(output_list: List[Any], input_list: List[Any], src: int = 0):
        dist.scatter_object_list(output_list, input_list, src)

    @staticmethod
    def gather_any(obj, obj_list: List[Any], dst: int = 0):
        dist.gather_object(obj, obj_list, dst)

    @staticmethod
    def gather_all(obj: Any) -> List[Any]:
        """all_gather_any returning the list (convenience for the layout / p2p rendezvous)."""
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out

    # ---- p2p primitives of the gloo data path -----------------------------------------
    @staticmethod
    def sync_send(tensor: Tensor, dst: int, tag: MessageType):
        return dist.send(tensor, dst, tag=tag.value)

    @staticmethod
    def sync_recv(tensor: Tensor, src: int, tag: MessageType):
        return dist.recv(tensor, src, tag=tag.value)

    @staticmethod
    def async_send(tensor: Tensor, dst: int, tag: MessageType):
        return dist.isend(tensor, dst, tag=tag.value)

    @staticmethod
    def async_recv(tensor: Tensor, src: int, tag: MessageType):
        return dist.irecv(tensor, src, tag=tag.value)

    # ---- gloo exchanges (reference data path; baseline + CPU plumbing) ---------------
    def fp_msg_exchange(self, recv_buffer_cpu: Basic_Buffer_Type, recv_buffer_gpu: Basic_Buffer_Type,
                        send_buffer_cpu: Basic_Buffer_Type, send_idx: Dict[int, Tuple[int, int]],
                        send_messages: Tensor):
        """fp32 all-to-all over W-1 ring rounds through host buffers (comm.py:166-191)."""
        rank, W = self.get_rank(), self.get_world_size()
        sends, recvs = [], Queue()
        for step in range(1, W):
            dst, src = (rank + step) % W, (rank - step + W) % W
            lo, hi = send_idx[dst]
            send_buffer_cpu[dst].copy_(send_messages[lo:hi])
            sends.append(self.async_send(send_buffer_cpu[dst], dst, MessageType.DATA))
            recvs.put((self.async_recv(recv_buffer_cpu[src], src, MessageType.DATA), src))
        while not recvs.empty():
            req, src = recvs.get()
            req.wait()
            recv_buffer_gpu[src].copy_(recv_buffer_cpu[src], non_blocking=True)
        for req in sends:
            req.wait()

    def qt_msg_exchange(self, recv_buffer_cpu: Basic_Buffer_Type, recv_buffer_gpu: Basic_Buffer_Type,
                        send_buffer_cpu: Basic_Buffer_Type):
        """Quantised all-to-all: int8 stream (tag DATA) + bf16 params (tag PARAMs) per peer
        (comm.py:193-222)."""
        rank, W = self.get_rank(), self.get_world_size()
        sends, recvs = [], Queue()
        for step in range(1, W):
            dst, src = (rank + step) % W, (rank - step + W) % W
            q_data, q_params = send_buffer_cpu[dst]
            sends.append(self.async_send(q_data, dst, MessageType.DATA))
            sends.append(self.async_send(q_params, dst, MessageType.PARAMs))
            recvs.put((self.async_recv(recv_buffer_cpu[src][0], src, MessageType.DATA),
                       self.async_recv(recv_buffer_cpu[src][1], src, MessageType.PARAMs), src))
        while not recvs.empty():
            r0, r1, src = recvs.get()
            r0.wait()
            r1.wait()
            recv_buffer_gpu[src][0].copy_(recv_buffer_cpu[src][0], non_blocking=True)
            recv_buffer_gpu[src][1].copy_(recv_buffer_cpu[src][1], non_blocking=True)
        for req in sends:
            req.wait()

    # ---- buffer management ---------------------------------------------------------------
    def init_buffer(self, *args, **kwargs):
        self.comm_buffer = CommBuffer(*args, **kwargs, device=self.device, transport=self.transport)

    def update_buffer(self, *args, **kwargs):
        assert self.comm_buffer is not None, "please initialize the communication buffer first"
        self.comm_buffer._update(*args, **kwargs)

    def delete_buffer(self, *args, **kwargs):
        assert self.comm_buffer is not None, "please initialize the communication buffer first"
        self.comm_buffer._delete(*args, **kwargs)


for _public in ("local_rank", "device", "init_method", "backend"):
    setattr(Communicator, _public, property(lambda self, _f="_" + _public: getattr(self, _f)))
