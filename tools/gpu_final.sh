#!/bin/bash
# round-end style validation on one B200: the GPU test suite, smoke(), both bench arms, the ncu launch list of the
# bench command and one full capture of the dominant (blur+composite) kernel.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (reference arm)"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref_err.txt; cut -c1-400 gpurun_out/final_bench_reference.json
echo "== bench"; timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench_err.txt; cut -c1-300 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench_err.txt
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/final_launches.csv
echo "== ncu full: k_post_fast"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_post_fast -s 2 -c 1 -f -o gpurun_out/final_prof_post \
    python bench.py --steps 2 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1; ls -la gpurun_out/final_prof_post.ncu-rep
