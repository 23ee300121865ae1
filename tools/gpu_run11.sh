mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pointwise" --csv --log-file gpurun_out/tcprobe_times.csv python tools/tc_probe.py > gpurun_out/tcprobe.log 2>&1
grep -E "k_pointwise" gpurun_out/tcprobe_times.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -10
for tc in "" "--tensor-cores"; do
python bench.py --workload deeplab720 --steps 5 --warmup 3 --streams 2 --batch 16 --no-cpu-baseline --no-e2e $tc | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['config']['pointwise_convs'][:16], 'value %.0f'%d['value'], 'cnn_us %.1f'%(1e3*d['stages']['cnn_ms_per_frame']))"
done
