#!/bin/bash
# DRAM traffic of the tcgen05 GEMM per tile rasterisation (ncu), 8192x4096x4096 products
mkdir -p gpurun_out
for gm in 1 8 -8 4; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:gemm_tf32x3 -s 9 -c 3 --csv --log-file gpurun_out/traffic_gm_$gm.csv \
      python scripts/gemm_bench.py --cg 2 --ksplit 0 --group-m $gm > /dev/null 2>&1
done
ls -la gpurun_out/traffic_gm_*
