mkdir -p gpurun_out
# launch list of one full-config bench step (per-launch device times; serialised, cold cache: compare shares)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:sketch_kernel|cmp|c16|fss_|fill_|RadixSort|densify' -c 300 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-genomes 64 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
wc -l gpurun_out/launches_r1b.csv
