#!/bin/bash
# session 2 baseline: whole GPU suite, the flat BagMinHash kernel (timing + ncu), bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest.txt
timeout 300 python scripts/sketch_only_bench.py 8 20000000 3 bmh 8192 -1 > gpurun_out/r2h_weighted.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 8 20000000 3 pmh 8192 -1 >> gpurun_out/r2h_weighted.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bmh_kernel -c 1 -o gpurun_out/r2h_bmh python scripts/sketch_only_bench.py 8 20000000 1 bmh 8192 -1 > gpurun_out/r2h_ncu_bmh.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -6 gpurun_out/r2h_pytest.txt; cat gpurun_out/r2h_weighted.txt; tail -3 gpurun_out/r2h_bench.err
