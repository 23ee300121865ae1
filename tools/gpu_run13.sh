mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/launches_warm_b64.csv \
   python bench.py --steps 2 --warmup 1 --streams 1 --batch 64 --no-e2e --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_warm_b64.csv
