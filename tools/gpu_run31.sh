mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "test_yuyv_ingest or test_app_stage_functions" > gpurun_out/racecheck_run31.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_run31.txt
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "test_yuyv_ingest or test_app_stage_functions or ragged" > gpurun_out/memcheck_run31.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_run31.txt
true
