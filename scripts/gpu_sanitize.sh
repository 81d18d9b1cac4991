mkdir -p gpurun_out
export D2G_FSS_BOOT_STRIDE=2
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest tests -m gpu -x -q -k "test_compare_codes_blocked_jobs and symmetric and 256 or test_compare_codes_special_values or test_sketch_chunked_upload or test_fss_guessed_bound or test_topk_matches_reference_golden or test_compressed_compare_golden and 1-False and codes" 2>&1 | tail -4
echo "memcheck rc=$?"; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
