#!/bin/bash
# round-2 two-GPU batch: cross-GPU parity tests, N=2 bench with the verify leg, NVLink timing + ncu link counters
set -u
O=gpurun_out/r2d; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -s -W ignore 2>&1 | grep -v UserWarning | grep -v "return torch.sparse" | tail -60) > $O/t_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $O/bench_n2.log 2> $O/bench_n2.err
timeout 500 python tools/bench_exchange.py --world 2 --devices 2 --scale 0.5 --reps 10 --json $O/exch_2gpu.json > $O/exch_2gpu.log 2>&1
M=nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,syslts__t_sectors_aperture_peer_op_write.sum,syslts__t_sectors_aperture_peer_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'send_quant|recv_quant|send_fp32' -c 16 --csv --log-file $O/ncu_nvlink_quant.csv python tools/bench_exchange.py --world 2 --devices 2 --scale 0.5 --reps 1 --only forward1:mixed > $O/ncu_quant.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'send_quant|recv_quant|send_fp32' -c 8 --csv --log-file $O/ncu_nvlink_fp32.csv python tools/bench_exchange.py --world 2 --devices 2 --scale 0.5 --reps 1 --only forward1:fp32 > $O/ncu_fp32.log 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sweep_overlap.py --steps 6 --json $O/sweep_n2.json > $O/sweep_n2.log 2>&1
tail -c 2500 $O/t_multi.log; tail -c 1500 $O/bench_n2.log; tail -3 $O/bench_n2.err; tail -12 $O/exch_2gpu.log; tail -5 $O/ncu_quant.log; tail -12 $O/sweep_n2.log
