mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "count_sketch or countsketch or weighted or multiset_and_prob or fail_loudly" > gpurun_out/pytest_cs.log 2>&1; tail -25 gpurun_out/pytest_cs.log
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_r1f.log python -m pytest tests -m gpu -x -q -k "test_distinct_kmers_matches_oracle_seeded and 21-30 or test_opmh_count_threshold_matches_reference_golden and k31 or test_count_sketch_weighted_matches_reference_golden and cs300" 2>&1 | tail -3
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_r1f.log
