mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "gpus_option" > gpurun_out/pytest_gpus2.log 2>&1; tail -5 gpurun_out/pytest_gpus2.log
timeout 400 python scripts/cli_vs_reference.py 256 5000000 1,2 > gpurun_out/cli_vs_ref_gpus.log 2>&1; tail -8 gpurun_out/cli_vs_ref_gpus.log
