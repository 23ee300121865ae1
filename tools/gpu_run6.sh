python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench.json
python bench.py --steps 10 --warmup 3 --streams 8 --batch 32 --no-cpu-baseline | tee gpurun_out/bench_s8.json
for wl in deeplab720 mlkit480 mlkit720; do python bench.py --workload $wl --steps 5 --warmup 3 --streams 2 --batch 16 --no-cpu-baseline --no-e2e | tee gpurun_out/bench_$wl.json; done
