mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pointwise_tile -c 2 -o gpurun_out/ncu_pw_tile -f python tools/ncu_pw.py tile 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gauss_fused -c 1 -o gpurun_out/ncu_gauss_fused -f python tools/ncu_pw.py gauss 2>&1 | tail -3
for wl in deeplab720; do
  for v in 2 0; do
    echo "== $wl variant $v"
    BSB_PW_VARIANT=$v timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_run18_${wl}_v$v.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
    python -c "
import json,sys
d=json.load(open('gpurun_out/bench_run18_${wl}_v$v.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
  done
done
echo "== camera blur 25 (fused gaussian, staged loads)"
timeout 300 python bench.py --bgblur 25 --camera-blur --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_run18_meet720_camblur25.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_run18_meet720_camblur25.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
ls -la gpurun_out/*.ncu-rep
