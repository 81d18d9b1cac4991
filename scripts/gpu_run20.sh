mkdir -p gpurun_out
python -m pytest tests/test_gpu_cli.py -m gpu -q > gpurun_out/pytest_cli.log 2>&1; tail -5 gpurun_out/pytest_cli.log
timeout 300 python scripts/cli_phases.py 256 5000000 > gpurun_out/cli_phases2.log 2>&1; grep -E "^==|pipelined|init|compare|ready" gpurun_out/cli_phases2.log | tail -40
timeout 300 python scripts/cli_vs_reference.py 256 5000000 > gpurun_out/cli_vs_ref2.log 2>&1; tail -4 gpurun_out/cli_vs_ref2.log
