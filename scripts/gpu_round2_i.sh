#!/bin/bash
# element streams: parity (library + front-end), then memcheck over a few of the new kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_streams.py -x -q 2>&1 | tail -25 > gpurun_out/r2i_pytest_streams.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -k "element_stream or protein or unsupported" 2>&1 | tail -25 > gpurun_out/r2i_pytest_cli.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "unsupported or parse_by_seq" 2>&1 | tail -8 > gpurun_out/r2i_pytest_parity.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_streams.py -x -q -k "k33 or k32_w33 or prot20_opmh_k5 or other_sketches" > gpurun_out/r2i_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2i_memcheck.txt
cat gpurun_out/r2i_pytest_streams.txt gpurun_out/r2i_pytest_cli.txt gpurun_out/r2i_pytest_parity.txt; tail -5 gpurun_out/r2i_memcheck.txt
