"""Per-epoch train/val/test metric store (reference: AdaQP/util/recorder.py:8-39)."""
from __future__ import annotations

import logging
import time
from typing import Sequence

import torch

logger = logging.getLogger("trainer")


class Recorder(object):
    def __init__(self, epoches: int):
        self.epoches_metrics = torch.zeros(epoches, 3)

    def add_new_metrics(self, epoch_count: int, epoch_metrics: Sequence[float]):
        """epoch_count runs from 1 to epoches; metrics = [train, val, test]."""
        assert len(epoch_metrics) == 3
        self.epoches_metrics[epoch_count - 1] = torch.tensor([float(m) for m in epoch_metrics])

    def summary(self):
        pct = 100 * self.epoches_metrics
        best = int(pct[:, 1].argmax())
        return {"Highest Train": float(pct[:, 0].max()), "Highest Valid": float(pct[:, 1].max()),
                "  Final Train": float(pct[best, 0]), "  Final Valid": float(pct[best, 1]),
                "   Final Test": float(pct[best, 2])}

    def display_final_statistics(self, metrics_file: str = None, val_metric_curve_file: str = None,
                                 model_name: str = "gcn"):
        rows = [f"{k}: {v:.2f}" for k, v in self.summary().items()]
        logger.info("\n" + "\n".join(rows))
        if metrics_file is not None:
            with open(metrics_file, "a") as f:
                f.write(f"{model_name} runs on {time.strftime('%Y-%m-%d', time.localtime())}:\n")
                f.write("\n".join(rows) + "\n")
        if val_metric_curve_file is not None:
            torch.save(100 * self.epoches_metrics[:, 1], val_metric_curve_file)
