#!/bin/bash
# round-2 second eight-GPU batch: all 8-GPU BASELINE configurations again, now with the tcgen05 dense GEMMs and the
# cost model fitted on the real kernel pair (adaptive); both arms, parity leg on
set -u
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/run_configs.py \
  --configs products-gcn-random,products-sage-adaptive,yelp-gcn-adaptive,amazon-sage-adaptive --steps 8 --ref-steps 3 --verify --out $O/r02b > $O/configs.log 2> $O/configs.err
grep '^{' $O/configs.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'error' in d: print(d['config_name'], d['impl'], 'ERROR', d['error'][-400:]); continue
    pc = d.get('parity_check') or {}
    print(d['config_name'], d.get('impl', 'ours'), round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3), 'exposed', round(d['exposed_comm_ms'], 2), d.get('assigned_bits_share_rank0'), d.get('cost_model_rank0'), 'mism', pc.get('mismatches'), 'act', (pc.get('activations') or {}).get('act_mean_rel_err'), (pc.get('activations') or {}).get('within_tolerance'))
"
tail -3 $O/configs.err
