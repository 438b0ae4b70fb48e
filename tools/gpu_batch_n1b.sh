#!/bin/bash
# round-2 one-GPU batch: accuracy experiment (8 ranks on one GPU), ncu --set full of the aggregation (v1 and v3),
# launch list of the N=1 bench, full GPU test-suite
set -u
O=gpurun_out/r2g; mkdir -p $O
(timeout 900 python tools/accuracy_check.py --spawn 8 --scale 0.05 --epochs 80 --seeds 3 --calibrate 0.03,0.05,0.08 --json $O/accuracy_gcn_w8.json 2>&1 | grep '^{' | cut -c1-400) > $O/accuracy.log
KM=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
for IMPL in 1 3; do
  ADAQP_SPMM=$IMPL timeout 500 ncu --set full --import-source on --clock-control none -k regex:spmm_csr -c 5 -f -o $O/spmm_n1_impl$IMPL python bench.py --steps 1 --warmup 3 --no-verify --no-cpu-baseline > $O/ncu_spmm_impl$IMPL.log 2>&1
  ncu -i $O/spmm_n1_impl$IMPL.ncu-rep --page raw --csv > $O/spmm_n1_impl${IMPL}_raw.csv 2>/dev/null
  ncu -i $O/spmm_n1_impl$IMPL.ncu-rep --page raw --csv --metrics $KM > $O/spmm_n1_impl${IMPL}_key.csv 2>/dev/null
done
ls -la $O/*.ncu-rep
# keep the reports only if they fit the 64 MiB return budget
SZ=$(du -cm $O/*.ncu-rep | tail -1 | cut -f1); if [ "$SZ" -gt 40 ]; then rm -f $O/spmm_n1_impl3.ncu-rep; fi
SZ=$(du -cm $O/*.ncu-rep 2>/dev/null | tail -1 | cut -f1); if [ "${SZ:-0}" -gt 40 ]; then rm -f $O/*.ncu-rep; fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_n1.csv python bench.py --steps 2 --warmup 3 --no-verify --no-cpu-baseline > $O/ncu_launches.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q -x -W ignore 2>&1 | tail -8) > $O/t_all.log
cat $O/accuracy.log | tail -30; tail -4 $O/t_all.log; head -c 600 $O/spmm_n1_impl1_key.csv
