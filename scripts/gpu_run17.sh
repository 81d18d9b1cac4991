mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "gpus_option or count_threshold_matches_oracle" > gpurun_out/pytest_gpus.log 2>&1; tail -30 gpurun_out/pytest_gpus.log
