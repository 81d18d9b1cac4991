#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fast_windowed or packed_entry or seeded or golden or long_walk or guessed or read_set" 2>&1 | tail -15 > gpurun_out/r2c_pytest_new.txt
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 3 fss 4096 51 > gpurun_out/r2c_sketch_only.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 >> gpurun_out/r2c_sketch_only.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 2 opmh 1024 -1 >> gpurun_out/r2c_sketch_only.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fast_kernel -c 1 -o gpurun_out/r2c_fast python scripts/sketch_only_bench.py 2048 5000000 1 fss 4096 51 > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_pytest_new.txt; cat gpurun_out/r2c_sketch_only.txt
