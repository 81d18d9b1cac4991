#!/bin/bash
# element streams + neighbour graphs: parity (library + front-end), memcheck over the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_streams.py tests/test_gpu_graphs.py -q --tb=short -rf 2>&1 | tail -60 > gpurun_out/r2j_pytest_new.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -q --tb=short -rf -k "element_stream or protein or unsupported or threshold or topk" 2>&1 | tail -40 > gpurun_out/r2j_pytest_cli.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "unsupported or parse_by_seq or topk" 2>&1 | tail -8 > gpurun_out/r2j_pytest_parity.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_streams.py tests/test_gpu_graphs.py -x -q -k "k33 or k32_w33 or prot20_opmh_k5 or other_sketches or (threshold and golden) or (fastcmp and golden)" > gpurun_out/r2j_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2j_memcheck.txt
grep -c . gpurun_out/r2j_pytest_new.txt; tail -15 gpurun_out/r2j_pytest_new.txt; tail -8 gpurun_out/r2j_pytest_cli.txt; tail -3 gpurun_out/r2j_pytest_parity.txt; tail -5 gpurun_out/r2j_memcheck.txt
