mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "parse_by_seq or distinct or many_small or unsupported" > gpurun_out/pytest_byseq.log 2>&1; tail -30 gpurun_out/pytest_byseq.log
