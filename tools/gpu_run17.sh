mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pointwise_variants or app_" 2>&1 | tail -3
timeout 300 python tools/pw_sweep.py 32 > gpurun_out/pw_sweep_b32.txt 2>&1; cat gpurun_out/pw_sweep_b32.txt
for wl in deeplab720 bodypix4k meet720; do
  for v in 2 0 3; do
    echo "== $wl variant $v"
    BSB_PW_VARIANT=$v timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_run17_${wl}_v$v.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
    python -c "
import json,sys
d=json.load(open('gpurun_out/bench_run17_${wl}_v$v.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
  done
done
echo "== camera blur 25 (fused gaussian)"
timeout 300 python bench.py --bgblur 25 --camera-blur --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_run17_meet720_camblur25.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_run17_meet720_camblur25.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
