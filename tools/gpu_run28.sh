mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for wl in meet720 mlkit480 deeplab720; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/bench_run28_$wl.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_run28_$wl.json')); print('$wl', round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/launches_warm_b64_run28.csv \
   python bench.py --steps 2 --warmup 1 --streams 1 --batch 64 --no-e2e --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_warm_b64_run28.csv
