mkdir -p gpurun_out
for wl in deeplab720 bodypix4k; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none -c 300 --csv --log-file gpurun_out/launches_warm_${wl}.csv \
   python bench.py --workload $wl --steps 1 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_warm_${wl}.csv
done
