#!/bin/bash
# two GPUs: the library's sharded comparison against the replicated one, the front-end with --gpus 2, bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_cmp_check.py 5000 4096 > gpurun_out/r2g_sharded.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/sharded_cmp_check.py 40000 1024 >> gpurun_out/r2g_sharded.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "gpus_option" 2>&1 | tail -5 > gpurun_out/r2g_pytest_cli.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cli > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
grep -v "^W\|^\[W\|warn" gpurun_out/r2g_sharded.txt | tail -8; cat gpurun_out/r2g_pytest_cli.txt; tail -3 gpurun_out/r2g_bench_n2.err; head -c 400 gpurun_out/r2g_bench_n2.json
