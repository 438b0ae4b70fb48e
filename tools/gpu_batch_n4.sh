#!/bin/bash
# BASELINE config 2: Reddit GCN, 4 partitions, AdaQP uniform 4-bit, 4x B200, both arms, parity leg on
set -u
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525 tools/run_configs.py \
  --configs reddit-gcn-uniform4 --steps 8 --ref-steps 3 --verify --out $O/r02 > $O/configs.log 2> $O/configs.err
grep '^{' $O/configs.log | cut -c1-1500; tail -3 $O/configs.err
