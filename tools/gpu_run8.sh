mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5
for tc in "" "--tensor-cores"; do
python bench.py --workload deeplab720 --steps 5 --warmup 3 --streams 2 --batch 16 --no-cpu-baseline --no-e2e $tc | tee gpurun_out/bench_deeplab720_tc${tc:+1}.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_deeplab_tc.csv \
   python bench.py --workload deeplab720 --steps 2 --warmup 1 --streams 1 --batch 16 --no-e2e --no-cpu-baseline --tensor-cores > /dev/null 2>&1
