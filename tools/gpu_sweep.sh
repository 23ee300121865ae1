#!/bin/bash
# tools/gpu_sweep.sh — streams x batch sweep of the device-resident throughput (bench.py, no e2e / cpu legs)
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for wl in ${WORKLOADS:-meet720}; do
for s in ${STREAMS:-1 2 4 8}; do for b in ${BATCHES:-8 16 32}; do
  timeout 120 python bench.py --workload $wl --streams $s --batch $b --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'wl':'$wl','S':$s,'B':$b,'fps':round(d['value']),'post_frac':round(d['roofline']['frac'],3),'cnn_us':round(1e3*d['stages']['cnn_ms_per_frame'],2),'post_us':round(1e3*d['stages']['post_ms_per_frame'],2),'all_us':round(1e3*d['stages']['all_ms_per_frame'],2)}))" | tee -a gpurun_out/sweep.jsonl
done; done; done
