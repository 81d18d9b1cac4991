#!/bin/bash
# GPU session: new fast-kernel tests first, then the whole suite, sketch-only timing, ncu of the fast kernel, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fast_windowed or packed_entry or topk_large" 2>&1 | tail -15 > gpurun_out/r2b_pytest_new.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.txt
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 3 fss 4096 51 > gpurun_out/r2b_sketch_only.txt 2>&1
D2G_NO_FAST=1 timeout 300 python scripts/sketch_only_bench.py 2048 5000000 2 fss 4096 51 >> gpurun_out/r2b_sketch_only.txt 2>&1
D2G_DEBUG=1 timeout 300 python scripts/sketch_only_bench.py 256 5000000 1 fss 4096 51 >> gpurun_out/r2b_sketch_only.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 >> gpurun_out/r2b_sketch_only.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fast_kernel -c 1 -o gpurun_out/r2b_fast python scripts/sketch_only_bench.py 2048 5000000 1 fss 4096 51 > gpurun_out/r2b_ncu.log 2>&1
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_pytest_new.txt; tail -3 gpurun_out/r2b_pytest.txt; cat gpurun_out/r2b_sketch_only.txt
