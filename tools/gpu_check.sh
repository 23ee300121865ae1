#!/bin/bash
# tools/gpu_check.sh — what we run on the B200 box through gpurun: parity tests, smoke,
# a short bench, the ncu launch list and one full ncu capture of the post kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
if [ "${1:-}" = "ncu" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  tail -3 gpurun_out/ncu_list.log
  echo "== ncu full: k_post"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_post -s 2 -c 1 -f -o gpurun_out/prof_post \
      python bench.py --steps 2 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post.log 2>&1
  tail -2 gpurun_out/ncu_post.log
  echo "== ncu full: CNN kernels"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pointwise_rows|k_depthwise_strip|k_conv_direct|k_tconv2x2|k_bilateral" -s 40 -c 12 -f -o gpurun_out/prof_cnn \
      python bench.py --steps 2 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_cnn.log 2>&1
  tail -2 gpurun_out/ncu_cnn.log
fi
ls -la gpurun_out
