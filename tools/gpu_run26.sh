mkdir -p gpurun_out
# memcheck: one small pipeline per model family + the option paths (slow under the sanitizer: keep it to a few tests)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_pipeline_ragged_geometry or test_app_options or test_yuyv_ingest or test_pointwise_variants or test_app_stage_functions" > gpurun_out/memcheck_run26.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck_run26.txt
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_run26.txt
