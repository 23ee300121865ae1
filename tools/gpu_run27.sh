mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q > gpurun_out/memcheck_run27.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_run27.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_pipeline_bit_exact or test_app_options or test_yuyv_ingest or test_pointwise_variants or test_app_stage_functions or atrous" > gpurun_out/racecheck_run27.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_run27.txt
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_pipeline_bit_exact or test_app_options" > gpurun_out/initcheck_run27.txt 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/initcheck_run27.txt
true
