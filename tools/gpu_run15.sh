# round-1 validation after the bgblur / flip / vcam-resize / animated-ring work
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" ; nvidia-smi topo -m 2>/dev/null | head -14
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_run15.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_run15.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_run15_meet720.json 2> gpurun_out/bench_run15_meet720.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_run15_meet720.json
timeout 600 python bench.py --workload bodypix4k --steps 10 --warmup 3 > gpurun_out/bench_run15_bodypix4k.json 2> gpurun_out/bench_run15_bodypix4k.err; echo "bench4k rc=$?"; cut -c1-1200 gpurun_out/bench_run15_bodypix4k.json; tail -2 gpurun_out/bench_run15_bodypix4k.err
timeout 300 python bench.py --bgblur 25 --camera-blur --no-cpu-baseline --steps 10 > gpurun_out/bench_run15_meet720_camblur25.json 2> gpurun_out/bench_run15_camblur.err; echo "benchblur rc=$?"; cut -c1-700 gpurun_out/bench_run15_meet720_camblur25.json; tail -2 gpurun_out/bench_run15_camblur.err
