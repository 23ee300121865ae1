mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_pointwise" -c 8 -f -o gpurun_out/prof_tcprobe python tools/tc_probe.py > gpurun_out/tcprobe.log 2>&1
tail -3 gpurun_out/tcprobe.log
