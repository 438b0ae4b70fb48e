#!/bin/bash
# ncu --set full of the two tcgen05 GEMM kernels (CSV export) + smoke()
set -u
O=gpurun_out/r2p; mkdir -p $O
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 200 python -m oracle.make_golden --half $O/golden_half > $O/golden_half.log 2>&1; tail -3 $O/golden_half.log
(timeout 300 python -m pytest tests/test_gpu_codec.py tests/test_gpu_gemm.py -q -W ignore 2>&1 | tail -12 | cut -c1-300) > $O/t_codec.log; cat $O/t_codec.log
timeout 200 python tools/bench_gemm.py $O/gemm_kb.json 2>&1 | head -6 | cut -c1-400
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'gemm_tf32x3|wgrad_tf32x3' -c 6 -f -o $O/gemm_n1 python bench.py --steps 1 --warmup 3 --no-verify --no-cpu-baseline > $O/ncu_gemm.log 2>&1
ncu -i $O/gemm_n1.ncu-rep --page raw --csv > $O/gemm_n1_full_raw.csv 2>/dev/null
KM=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_op_hmma.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic
ncu -i $O/gemm_n1.ncu-rep --page raw --csv --metrics $KM > $O/gemm_n1_key.csv 2>/dev/null
rm -f $O/gemm_n1.ncu-rep
cut -c1-60,150-700 $O/gemm_n1_key.csv | head -10
ncu -i /dev/null 2>/dev/null; grep -i -E "tensor|utc" $O/gemm_n1_full_raw.csv | head -0
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2p/gemm_n1_full_raw.csv")))
h=rows[0]
for name in h:
    if "tensor" in name.lower() or "tmem" in name.lower() or "utc" in name.lower():
        i=h.index(name); print(name, [r[i] for r in rows[2:5]])
PY
