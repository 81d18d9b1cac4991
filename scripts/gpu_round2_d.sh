#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.txt
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 3 fss 4096 51 > gpurun_out/r2d_sketch_only.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 >> gpurun_out/r2d_sketch_only.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fast_kernel -c 1 -o gpurun_out/r2d_fast python scripts/sketch_only_bench.py 2048 5000000 1 fss 4096 51 > gpurun_out/r2d_ncu.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -5 gpurun_out/r2d_pytest.txt; cat gpurun_out/r2d_sketch_only.txt; tail -3 gpurun_out/r2d_bench.err; head -c 300 gpurun_out/r2d_bench.json
