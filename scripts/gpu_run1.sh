set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python scripts/cmp_only_bench.py 10000 4096 3 f64 > gpurun_out/cmpbench.log 2>&1
python scripts/cmp_only_bench.py 10000 4096 3 codes >> gpurun_out/cmpbench.log 2>&1
python scripts/cmp_only_bench.py 10000 4096 2 codes containment >> gpurun_out/cmpbench.log 2>&1
python scripts/cmp_only_bench.py 10000 4096 2 f64 containment >> gpurun_out/cmpbench.log 2>&1
python scripts/cmp_only_bench.py 20000 1024 3 codes >> gpurun_out/cmpbench.log 2>&1
python scripts/cmp_only_bench.py 20000 1024 3 f64 >> gpurun_out/cmpbench.log 2>&1
cat gpurun_out/cmpbench.log
