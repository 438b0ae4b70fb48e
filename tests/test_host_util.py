"""Host-side utilities that define the reported numbers (no GPU)."""
import time

import pytest
import torch

from adaqp_b200.util import Recorder, Timer


def test_timer_buckets_follow_reference_names():
    t = Timer(device=torch.device("cpu"))
    for name in ("forward0_communication", "forward0_quantization", "forward0_de-quantization",
                 "forward0_central_aggregation", "forward0_marginal_aggregation", "backward1_full_aggregation"):
        with t.record(name):
            time.sleep(0.002)
    comm, quant, central, marginal, full = t.epoch_traced_time()
    assert comm > 0 and central > 0 and marginal > 0 and full > 0
    assert quant >= 0.004 * 0.9            # quantization + de-quantization share one column (timer.py:51)
    with pytest.raises(Exception):
        with t.record("forward0_communication"):
            pass
    with pytest.raises(KeyError):
        with t.record("forward0_unknown"):
            pass
    t.clear()
    assert t.epoch_traced_time() == [0.0] * 5 and len(t._total_record) == 1


def test_recorder_summary():
    r = Recorder(3)
    r.add_new_metrics(1, [0.5, 0.4, 0.3])
    r.add_new_metrics(2, [0.7, 0.9, 0.8])
    r.add_new_metrics(3, [0.9, 0.6, 0.5])
    s = r.summary()
    assert s["Highest Valid"] == pytest.approx(90.0) and s["   Final Test"] == pytest.approx(80.0)


def test_mode_table_matches_reference():
    from adaqp_b200.trainer.trainer import QUNAT_PARA_MAP, RUNING_MODE
    assert RUNING_MODE == ["Vanilla", "AdaQP", "AdaQP-q", "AdaQP-p"]
    assert QUNAT_PARA_MAP == {"Vanilla": ("full", False), "AdaQP": ("quant", True), "AdaQP-q": ("quant", False),
                              "AdaQP-p": ("full", True)}


def test_alias_package_and_shims_import():
    import AdaQP
    import quant_cuda
    from AdaQP.communicator import BITS_SET, Communicator  # noqa: F401
    from AdaQP.helper import BitType, MessageType
    assert BITS_SET == (2, 4, 8) and MessageType.PARAMs.value == 1 and BitType.QUANT.value == 1
    assert hasattr(quant_cuda, "pack_single_precision") and hasattr(quant_cuda, "unpack_single_precision")
    assert AdaQP.Trainer.__name__ == "Trainer"
    with pytest.raises(NotImplementedError):
        Communicator(backend="nccl")
