#!/bin/bash
# round-2 final one-GPU batch on the final tree: launch list of the N=1 bench (share of the step per kernel), ncu --set full
# of the fused LayerNorm + ReLU kernels, then the plain N=1 bench line (not under a profiler).
O=gpurun_out/r2f; mkdir -p $O
KM=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_n1_final.csv python bench.py --steps 2 --warmup 3 --no-verify --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:ln_relu -c 4 -f -o $O/norm_n1 python bench.py --steps 1 --warmup 3 --no-verify --no-cpu-baseline > $O/ncu_norm.log 2>&1
ncu -i $O/norm_n1.ncu-rep --page raw --csv --metrics $KM > $O/norm_n1_key.csv 2>/dev/null
ncu -i $O/norm_n1.ncu-rep --page raw --csv > $O/norm_n1_full_raw.csv 2>/dev/null
rm -f $O/norm_n1.ncu-rep
timeout 240 python bench.py > $O/bench_n1_final.json 2> $O/bench_n1_final.err
tail -c 600 $O/bench_n1_final.json
