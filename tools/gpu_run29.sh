mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or app_options or yuyv or large_batch or 4k" 2>&1 | tail -2
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  for wl in meet720 bodypix4k; do
  BSB_POST_WIDE=$1 BSB_POST_L1=$2 timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/b.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
  python -c "
import json
d=json.load(open('gpurun_out/b.json')); print('wide=$1 l1=$2 $wl', round(d['value']), 'post us', round(d['stages']['post_ms_per_frame']*1e3,3), 'frac', round(d['roofline']['frac'],3))"
  done
done
