mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "count_threshold or parse_by_seq or unsupported or fail_loudly" > gpurun_out/pytest_mincount.log 2>&1; tail -30 gpurun_out/pytest_mincount.log
