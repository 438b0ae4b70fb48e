"""Trainer: wires Communicator -> GraphEngine -> CommBuffer -> Assigner -> model and runs the
epoch loop.  Public surface of AdaQP/trainer/trainer.py:23-238 kept (`Trainer(args)`,
`.train() -> Tensor[8]`, `.save(records)`, the --mode table, the CSV columns), so the
reference's main.py drives it unchanged."""
from __future__ import annotations

import csv
import os
from argparse import Namespace
from typing import Dict, Tuple

import torch
import yaml
from torch import Tensor

from ..assigner import Assigner as assigner
from ..communicator import Communicator as comm
from ..helper import BitType, DistGNNType
from ..manager import GraphEngine as engine
from ..model import DistGCN, DistSAGE
from .runtime_util import (aggregate_accuracy, aggregate_F1, setup_logger, sync_model, sync_seed,
                           train_for_one_epoch, val_test)

RUNING_MODE = ["Vanilla", "AdaQP", "AdaQP-q", "AdaQP-p"]
# mode -> (message precision, overlap central aggregation with the exchange)
QUNAT_PARA_MAP: Dict[str, Tuple[str, bool]] = {"Vanilla": ("full", False), "AdaQP": ("quant", True),
                                               "AdaQP-q": ("quant", False), "AdaQP-p": ("full", True)}
MODEL_MAP: Dict[str, DistGNNType] = {"gcn": DistGNNType.DistGCN, "sage": DistGNNType.DistSAGE}


class Trainer(object):
    def __init__(self, runtime_args: Namespace):
        args = vars(runtime_args) if not isinstance(runtime_args, dict) else dict(runtime_args)
        dataset = args["dataset"]
        cfg_path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "config", f"{dataset}.yaml")
        with open(cfg_path, "r") as f:
            self.config = yaml.load(f, Loader=yaml.FullLoader)
        self.config["runtime"].update({k: v for k, v in args.items() if v is not None})
        if os.environ.get("ADAQP_NUM_EPOCHES"):          # short runs of the unmodified reference main.py (no such flag there)
            self.config["runtime"]["num_epoches"] = int(os.environ["ADAQP_NUM_EPOCHES"])
        # extension: `assign_bits` / `assign_cycle` / `group_size` given at run time override the yaml's
        # `assignment:` section (the reference edits the yaml, e.g. assign_bits: 4 for uniform 4-bit)
        for k in ("assign_bits", "assign_cycle", "group_size", "coe_lambda"):
            if args.get(k) is not None:
                self.config["assignment"][k] = args[k]
        rt = self.config["runtime"]
        self.exp_path = f"{rt['exp_path']}/{dataset}/{rt['num_parts']}part/{rt['model_name']}"
        self.logger = setup_logger("trainer.log", rt["logger_level"], with_file=True)
        self._set_communicator()
        self._set_engine()
        if comm.get_rank() == 0:
            os.makedirs(self.exp_path, exist_ok=True)
        self._set_buffer()
        self._set_assigner()
        if engine.ctx.bit_type == BitType.QUANT:
            # adaptive starts from the uniform default until variances have been traced (:62-69)
            first = "uniform" if assigner.ctx.scheme == "adaptive" else None
            comm.ctx.update_buffer(assigner.ctx.get_assignment(engine.ctx.send_idx, runtime_scheme=first))
        self._set_model()

    # ---- setup --------------------------------------------------------------------------------
    def _set_communicator(self):
        rt = self.config["runtime"]
        self.communicator = comm(rt["backend"], rt["init_method"])
        self.logger.info(repr(self.communicator))

    def _set_engine(self):
        data, rt, model = self.config["data"], self.config["runtime"], self.config["model"]
        if rt["mode"] not in RUNING_MODE:
            raise ValueError(f"Invalid running mode: {rt['mode']}")
        if rt["model_name"] not in MODEL_MAP:
            raise ValueError(f"Invalid model type: {rt['model_name']}")
        precision, use_parallel = QUNAT_PARA_MAP[rt["mode"]]
        self.engine = engine(rt["num_epoches"], data["partition_path"], rt["dataset"], precision,
                             MODEL_MAP[rt["model_name"]], use_parallel)
        engine.ctx.agg_type = model["aggregator_type"]
        if engine.ctx.use_parallel:
            for g in (engine.ctx.graph, engine.ctx.bwd_graph):
                g.init_copy_buffers(data["num_feats"], model["hidden_dim"], model["num_layers"], engine.ctx.device)
        self.logger.info(repr(self.engine))

    def _set_buffer(self):
        data, model = self.config["data"], self.config["model"]
        shape = [data["num_feats"]] + [model["hidden_dim"]] * (model["num_layers"] - 1)
        comm.ctx.init_buffer(shape, engine.ctx.send_idx, engine.ctx.recv_idx, engine.ctx.bit_type,
                             total_send_idx=engine.ctx.total_send_idx, num_remote=engine.ctx.num_remove)

    def _set_assigner(self):
        data, model, rt, asg = (self.config[k] for k in ("data", "model", "runtime", "assignment"))
        self.assigner = assigner(data["num_feats"], model["hidden_dim"], model["num_layers"],
                                 asg["profile_data_length"], rt["assign_scheme"], asg["assign_bits"],
                                 engine.ctx.scores, asg["group_size"], asg["coe_lambda"], asg["assign_cycle"])
        self.logger.info(self.assigner)

    def _set_model(self):
        data, model, rt = self.config["data"], self.config["model"], self.config["runtime"]
        kind = MODEL_MAP[rt["model_name"]]
        common = (data["num_feats"], model["hidden_dim"], data["num_classes"], model["num_layers"],
                  model["dropout_rate"], model["use_norm"])
        if kind == DistGNNType.DistGCN:
            self.model = DistGCN(*common).to(comm.ctx.device)
        else:
            self.model = DistSAGE(*common, model["aggregator_type"]).to(comm.ctx.device)

    # ---- runtime ----------------------------------------------------------------------------------
    def train(self):
        rt = self.config["runtime"]
        multilabel = self.config["data"]["is_multilabel"]
        sync_seed()
        self.model.reset_parameters()
        sync_model(self.model)
        optimizer = torch.optim.Adam(self.model.parameters(), lr=rt["learning_rate"], weight_decay=rt["weight_decay"])
        criterion = torch.nn.BCEWithLogitsLoss(reduction="sum") if multilabel else torch.nn.CrossEntropyLoss(reduction="sum")
        eng = self.engine.ctx
        feats, labels = eng.feats, eng.labels
        n_train = torch.LongTensor([eng.train_mask.numel()])
        comm.all_reduce_sum(n_train)
        n_train = n_train.item()
        assign_time, train_time = [], []
        self.exposed_comm_ms = []
        for epoch in range(1, rt["num_epoches"] + 1):
            overhead, loss, traced, reduce_time = train_for_one_epoch(
                epoch, eng.graph, self.model, feats, labels, optimizer, criterion, n_train, eng.train_mask)
            assign_time.append(overhead)
            train_time.append(traced)
            self.exposed_comm_ms.append(getattr(eng, "last_exposed_comm_ms", 0.0))
            metrics = val_test(eng.graph, self.model, feats, labels, eng.train_mask, eng.val_mask, eng.test_mask, multilabel)
            info = (aggregate_F1 if multilabel else aggregate_accuracy)(loss, metrics, epoch)
            if epoch % rt["log_steps"] == 0:
                if comm.get_rank() == 0:
                    if not eng.use_parallel:
                        t = (f"Worker 0 | Total Time {traced[0]:.4f}s | Comm Time {traced[1]:.4f}s | Quant Time "
                             f"{traced[2]:.4f}s | Agg Time {traced[-1]:.4f}s | Reduce Time {reduce_time:.4f}s")
                    else:
                        t = (f"Worker 0 | Total Time {traced[0]:.4f}s | Comm Time {traced[1]:.4f}s | Quant Time "
                             f"{traced[2]:.4f}s | Central Agg Time {traced[3]:.4f}s | Marginal Agg Time "
                             f"{traced[4]:.4f}s | Reduce Time {reduce_time:.4f}s | Exposed Comm "
                             f"{self.exposed_comm_ms[-1]:.3f}ms")
                    self.logger.info(info + "\n" + t)
                comm.barrier()
        tt = torch.tensor(train_time)
        records = torch.concat([torch.tensor(assign_time).sum().view(-1), tt.sum(dim=0)[0].view(-1), tt.mean(dim=0)])
        comm.ctx.delete_buffer()
        return records

    def save(self, time_records: Tensor):
        if comm.get_rank() != 0:
            comm.gather_any(time_records, None, dst=0)
            comm.barrier()
            return
        rows = [None] * comm.get_world_size()
        comm.gather_any(time_records, rows, dst=0)
        paths = {k: f"{self.exp_path}/{k}" for k in ("metrics", "time", "val_curve")}
        for p in paths.values():
            os.makedirs(p, exist_ok=True)
        name = self.config["runtime"]["mode"]
        if engine.ctx.bit_type == BitType.QUANT:
            name = f"{name}_{self.config['runtime']['assign_scheme']}"
        engine.ctx.recorder.display_final_statistics(f"{paths['metrics']}/{name}.txt", f"{paths['val_curve']}/{name}.pt",
                                                     self.config["runtime"]["model_name"])
        csv_path = f"{paths['time']}/{name}.csv"
        new_file = not os.path.exists(csv_path)
        with open(csv_path, "a") as f:
            w = csv.writer(f)
            if new_file:
                w.writerow(["Worker", "Overhead", "Total", "Per_epoch", "Comm", "Quant", "Central", "Marginal", "Full"])
            for worker, rec in enumerate(rows):
                line = [f"Worker {worker}"] + list(rec.numpy())
                assert len(line) == 9, f"Invalid write data length: {len(line)}"
                w.writerow(line)
        comm.barrier()
