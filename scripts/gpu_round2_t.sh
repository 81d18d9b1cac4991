#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rf -k "weighted or bmh or multiset or config3 or other_sketches or kmercounts or ids_" 2>&1 | tail -5 > gpurun_out/r2t_pytest.txt; tail -3 gpurun_out/r2t_pytest.txt
timeout 300 python scripts/sketch_only_bench.py 8 20000000 3 bmh 8192 -1 > gpurun_out/r2t_weighted.txt 2>&1; tail -2 gpurun_out/r2t_weighted.txt
timeout 300 python scripts/sketch_only_bench.py 64 20000000 2 bmh 8192 -1 >> gpurun_out/r2t_weighted.txt 2>&1; tail -1 gpurun_out/r2t_weighted.txt
