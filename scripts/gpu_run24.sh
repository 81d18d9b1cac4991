mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "topk or test_compare_opmh_golden or test_gpus_option" > gpurun_out/pytest_nlsh.log 2>&1; tail -25 gpurun_out/pytest_nlsh.log
