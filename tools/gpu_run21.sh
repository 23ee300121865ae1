mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "atrous or every_tensor" 2>&1 | tail -2
for wl in deeplab720 bodypix4k; do
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_run21_${wl}.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
    python -c "
import json,sys
d=json.load(open('gpurun_out/bench_run21_${wl}.json')); print('$wl', round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
done
for wl in deeplab720 bodypix4k; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none -c 300 --csv --log-file gpurun_out/launches_warm_${wl}.csv \
   python bench.py --workload $wl --steps 1 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_warm_${wl}.csv
done
