"""Named region timer with the reference's bucket semantics (AdaQP/util/timer.py:18-51).

Region names must contain one of: communication, quantization, de-quantization, central,
marginal, full.  `record(name)` synchronises the current stream on entry and exit and
takes wall-clock time, exactly like the reference (its synchronisation is what makes the
reference's "Comm/Quant/Agg" columns comparable).  The B200 path adds `record_events`,
which brackets a region with CUDA events instead and therefore does not stall the host:
those durations are resolved lazily in `epoch_traced_time`.
"""
from __future__ import annotations

import time
from contextlib import contextmanager

import torch

BUCKETS = ("communication", "quantization", "de-quantization", "central", "marginal", "full")


def _bucket(name: str) -> int:
    if "communication" in name:
        return 0
    if "quantization" in name and "de" not in name:
        return 1
    if "de-quantization" in name:
        return 2
    if "central" in name:
        return 3
    if "marginal" in name:
        return 4
    if "full" in name:
        return 5
    raise KeyError(f"no {name} key")


class Timer(object):
    def __init__(self, device: torch.device):
        self._device = device
        self._record = {}
        self._events = {}
        self._exposed = []
        self._exposed = []
        self._total_record = []
        self.use_cuda = device is not None and torch.device(device).type == "cuda"

    def _sync(self):
        if self.use_cuda:
            torch.cuda.current_stream(self._device).synchronize()

    @contextmanager
    def record(self, name: str):
        if name in self._record or name in self._events:
            raise Exception(f"{name} already exists")
        _bucket(name)
        self._sync()
        start = time.time()
        yield
        self._sync()
        self._record[name] = (start, time.time())

    @contextmanager
    def record_events(self, name: str, stream=None):
        """Device-side timing of a region on `stream` without host synchronisation."""
        if name in self._record or name in self._events:
            raise Exception(f"{name} already exists")
        _bucket(name)
        if not self.use_cuda:
            start = time.time()
            yield
            self._record[name] = (start, time.time())
            return
        s = stream if stream is not None else torch.cuda.current_stream(self._device)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record(s)
        yield
        b.record(s)
        self._events[name] = (a, b)

    def record_exposed(self, name: str, compute_done, data_ready):
        """Exposed communication of one layer pass: how long the default stream had to wait
        for remote data after its overlappable work finished = max(0, t(data_ready) -
        t(compute_done)) (the `response.get()` wait of ops.py:177)."""
        self._exposed.append((name, compute_done, data_ready))

    def exposed_comm_ms(self) -> float:
        if not self._exposed:
            return 0.0
        torch.cuda.synchronize(self._device)
        return float(sum(max(0.0, a.elapsed_time(b)) for _, a, b in self._exposed))

    def epoch_traced_time(self):
        tot = [0.0] * 6
        for name, (start, end) in self._record.items():
            tot[_bucket(name)] += end - start
        if self._events:
            torch.cuda.synchronize(self._device)
            for name, (a, b) in self._events.items():
                tot[_bucket(name)] += a.elapsed_time(b) / 1e3
        # [comm, quant + dequant, central, marginal, full]  (timer.py:51)
        return [tot[0], tot[1] + tot[2], tot[3], tot[4], tot[5]]

    def clear(self, is_train: bool = True):
        if is_train:
            self._total_record.append(self.epoch_traced_time())
        self._record = {}
        self._events = {}
        self._exposed = []
