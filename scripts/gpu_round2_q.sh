#!/bin/bash
# launch list of the headline bench (library kernels only) with DRAM bytes per launch
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:sketch|pack_ascii|fss_|cmp16|c16_|fill_u64|cmp_tile|DeviceSegmented|DeviceRadix' --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cli --no-verify --e2e-genomes 64 > gpurun_out/r2q_bench_under_ncu.log 2>&1
grep -c . gpurun_out/r2q_launches.csv; tail -2 gpurun_out/r2q_bench_under_ncu.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_streams.py -q -k "ids_and_counts" 2>&1 | tail -3
