mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "save_kmers or fss" > gpurun_out/pytest_fssids.log 2>&1; tail -25 gpurun_out/pytest_fssids.log
