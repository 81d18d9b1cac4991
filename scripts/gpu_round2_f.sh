#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest_cli.txt
timeout 600 python scripts/e2e_sketch_bench.py 1024 5000000 > gpurun_out/r2f_e2e.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 8 20000000 2 bmh 8192 -1 > gpurun_out/r2f_weighted.txt 2>&1
timeout 300 python scripts/sketch_only_bench.py 8 20000000 2 pmh 8192 -1 >> gpurun_out/r2f_weighted.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bmh_kernel -c 1 -o gpurun_out/r2f_bmh python scripts/sketch_only_bench.py 8 20000000 1 bmh 8192 -1 > gpurun_out/r2f_ncu_bmh.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pmh_kernel -c 1 -o gpurun_out/r2f_pmh python scripts/sketch_only_bench.py 8 20000000 1 pmh 8192 -1 > gpurun_out/r2f_ncu_pmh.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_bmh.csv python scripts/sketch_only_bench.py 8 20000000 1 bmh 8192 -1 > /dev/null 2>&1
tail -4 gpurun_out/r2f_pytest_cli.txt; cat gpurun_out/r2f_e2e.txt gpurun_out/r2f_weighted.txt
