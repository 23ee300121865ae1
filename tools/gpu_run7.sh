mkdir -p gpurun_out
for wl in deeplab720; do for tc in "" "--tensor-cores"; do
python bench.py --workload $wl --steps 5 --warmup 3 --streams 2 --batch 16 --no-cpu-baseline --no-e2e $tc | tee gpurun_out/bench_${wl}_tc${tc:+1}.json
done; done
python bench.py --workload deeplab720 --steps 5 --warmup 3 --streams 4 --batch 32 --no-cpu-baseline --no-e2e --tensor-cores | tee gpurun_out/bench_deeplab720_tc_s4b32.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pointwise_tc -s 20 -c 6 -f -o gpurun_out/prof_tc \
   python bench.py --workload deeplab720 --steps 2 --warmup 1 --streams 1 --batch 16 --no-e2e --no-cpu-baseline --tensor-cores > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_deeplab_tc.csv \
   python bench.py --workload deeplab720 --steps 2 --warmup 1 --streams 1 --batch 16 --no-e2e --no-cpu-baseline --tensor-cores > /dev/null 2>&1
