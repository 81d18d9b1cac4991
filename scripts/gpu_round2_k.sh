#!/bin/bash
# per-config bench lines (configs 1, 3, 4, 5) + ncu captures of their dominant kernels + CLI tests that failed last time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_graphs.py -q --tb=short -rf -k "protein or threshold" 2>&1 | tail -12 > gpurun_out/r2k_pytest.txt
for c in 1 4 5 3; do
  timeout 1500 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2k_bench_c$c.json 2> gpurun_out/r2k_bench_c$c.err
  echo "config $c rc=$?"; tail -c 600 gpurun_out/r2k_bench_c$c.err
done
# ncu: one capture per dominant kernel (reduced sizes; the kernels and their launch geometry per unit of work are the same)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -c 1 -o gpurun_out/r2k_opmh python bench.py --config 1 --steps 1 --warmup 1 > gpurun_out/r2k_ncu_opmh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lsh_refine_kernel -c 1 -o gpurun_out/r2k_lsh_refine python bench.py --config 5 --n 100000 --steps 1 --warmup 0 > gpurun_out/r2k_ncu_lsh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lsh_query_kernel -c 1 -o gpurun_out/r2k_lsh_query python bench.py --config 5 --n 100000 --steps 1 --warmup 0 >> gpurun_out/r2k_ncu_lsh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cmp16_tile_kernel -c 1 -o gpurun_out/r2k_panel python bench.py --config 4 --n 10000 --steps 1 --warmup 0 > gpurun_out/r2k_ncu_panel.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_c5.csv python bench.py --config 5 --n 100000 --steps 1 --warmup 0 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_c1.csv python bench.py --config 1 --steps 1 --warmup 1 > /dev/null 2>&1
cat gpurun_out/r2k_pytest.txt | tail -4
for c in 1 4 5 3; do head -c 500 gpurun_out/r2k_bench_c$c.json; echo; done
