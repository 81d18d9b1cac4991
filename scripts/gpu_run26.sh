mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "nlsh or test_topk_matches_reference_golden or test_cmp_topk_csr_file" > gpurun_out/pytest_nlsh3.log 2>&1; tail -15 gpurun_out/pytest_nlsh3.log
