mkdir -p gpurun_out
timeout 400 python tools/pw_sweep.py 32 > gpurun_out/pw_sweep3_b32.txt 2>&1; cat gpurun_out/pw_sweep3_b32.txt
timeout 300 python bench.py --workload mlkit480 > gpurun_out/bench_run22_mlkit480.json 2>gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_run22_mlkit480.json')); print('mlkit480', round(d['value']), round(d['e2e']['value']), d['cpu_baseline']['value'], d['roofline']['frac'], {k:round(v*1e3,2) for k,v in d['stages'].items()})"
