mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "save_kmers or weighted or count_sketch or multiset_and_prob" > gpurun_out/pytest_ids.log 2>&1; tail -25 gpurun_out/pytest_ids.log
