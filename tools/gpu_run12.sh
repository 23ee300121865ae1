mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --streams 2 --batch 32 --no-cpu-baseline | tee gpurun_out/bench_s2.json
python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench.json
python bench.py --workload mlkit480 --steps 10 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_mlkit480.json
