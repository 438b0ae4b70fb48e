"""adaqp_b200: B200-native boundary-message exchange + aggregation for AdaQP.

Host-side mirror of the reference's Python interface for the hot path (Trainer,
GraphEngine/DecompGraph, Communicator/CommBuffer, Assigner, quant_cuda) over the C ABI
of libadaqp_b200.so (include/adaqp_b200.h).  Sub-modules import lazily so that the
CPU-only control-plane pieces stay importable without a GPU.
"""
__all__ = ["Trainer"]


def __getattr__(name):
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    raise AttributeError(name)
