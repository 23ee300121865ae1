mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pointwise_variants or app_" 2>&1 | tail -2
python - <<'PY'
import backscrub_b200 as bs
lib = bs.lib()
for (M,K,N) in [(34848,512,256),(34848,256,256),(34848,480,160),(34848,160,256),(34848,128,256),(34848,480,80),(8712,256,256),(8712,128,256)]:
    t2 = lib.bsb_time_pointwise(0,2,M,K,N,20); t3 = lib.bsb_time_pointwise(0,3,M,K,N,20)
    print(M,K,N, f"old {t2:.4f} tile {t3:.4f} {t2/t3:.2f}x {2.0*M*K*N/t3/1e9:.1f} TF")
PY
for wl in deeplab720 bodypix4k; do
    BSB_PW_VARIANT=0 timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_run19_${wl}.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
    python -c "
import json,sys
d=json.load(open('gpurun_out/bench_run19_${wl}.json')); print('$wl', round(d['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
done
echo "== camera blur 25"
timeout 300 python bench.py --bgblur 25 --camera-blur --no-cpu-baseline --steps 10 > gpurun_out/bench_run19_meet720_camblur25.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_run19_meet720_camblur25.json')); print(round(d['value']), round(d['e2e']['value']), {k:round(v*1e3,2) for k,v in d['stages'].items()})"
