#!/bin/bash
# round-2 eight-GPU batch: BASELINE configs 3-5 + the headline config, both arms, parity leg on; SM-cap sweep
set -u
O=gpurun_out/r2f; mkdir -p $O profiles/bench
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/run_configs.py \
  --configs products-gcn-random,products-sage-adaptive,yelp-gcn-adaptive,amazon-sage-adaptive --steps 8 --ref-steps 3 --verify --out $O/r02 > $O/configs.log 2> $O/configs.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/sweep_overlap.py --steps 8 --caps 0:0,112:112,96:96,64:64,96:0,64:0,0:64 --json $O/sweep_n8.json > $O/sweep_n8.log 2>&1
grep '^{' $O/configs.log | cut -c1-600; tail -3 $O/configs.err; grep '^{' $O/sweep_n8.log
