#!/bin/bash
# whole GPU suite (filter set, per-list refine kernel included), config 5 / 3 lines again, ncu of the new refine kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -30 > gpurun_out/r2l_pytest.txt
timeout 1500 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r2l_bench_c5.json 2> gpurun_out/r2l_bench_c5.err; echo "c5 rc=$?"
D2G_LSH_REFINE_PER_ENTRY=1 timeout 900 python bench.py --config 5 --steps 2 --warmup 1 > gpurun_out/r2l_bench_c5_perentry.json 2>> gpurun_out/r2l_bench_c5.err
timeout 1500 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2l_bench_c3.json 2> gpurun_out/r2l_bench_c3.err; echo "c3 rc=$?"; tail -c 400 gpurun_out/r2l_bench_c3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lsh_refine_list_kernel -c 1 -o gpurun_out/r2l_lsh_refine_list python bench.py --config 5 --n 100000 --steps 1 --warmup 0 > gpurun_out/r2l_ncu_lsh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bmh_kernel -c 1 -o gpurun_out/r2l_bmh python bench.py --config 3 --genomes 8 --steps 1 --warmup 0 > gpurun_out/r2l_ncu_bmh.log 2>&1
tail -12 gpurun_out/r2l_pytest.txt
for f in r2l_bench_c5 r2l_bench_c5_perentry r2l_bench_c3; do head -c 700 gpurun_out/$f.json; echo; done
