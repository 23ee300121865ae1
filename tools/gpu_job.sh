#!/bin/bash
# tools/gpu_job.sh <tag> <part>... — the one script behind every `gpurun` call of this round.
# Parts (each bounded by its own timeout, results under gpurun_out/<tag>_*):
#   tests            pytest -m gpu
#   tests:<expr>     pytest -m gpu -k <expr>
#   bench:<name>:<args...>   python bench.py <args> > gpurun_out/<tag>_bench_<name>.json   (args with ',' for spaces)
#   launches:<name>:<args>   ncu launch list (gpu__time_duration) of a short bench run
#   ncu:<name>:<kernel-regex>:<args>   ncu --set full of one kernel -> .ncu-rep + summary
#   ncux:<name>:<kernel-regex>:<skip>:<count>:<args>   same with an explicit launch window (regex may use | ; quote it)
#   smoke            __graft_entry__.smoke()
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
for part in "$@"; do
  IFS=':' read -r kind name a3 a4 a5 a6 <<< "$part"
  case $kind in
    tests)
      if [ -n "${name:-}" ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$name" > gpurun_out/${TAG}_pytest_$(echo $name | tr ' ' '_').txt 2>&1
      else timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; fi
      tail -3 gpurun_out/${TAG}_pytest*.txt ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt ;;
    bench)
      args=$(echo "${a3:-}" | tr ',' ' ')
      timeout 900 python bench.py $args > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err || tail -5 gpurun_out/${TAG}_bench_${name}.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("${name}: value %.0f e2e %s post %.1f us/launch frac %.3f cnn %.2f us/frame launches %s" % (
        d["value"], (d.get("e2e") or {}).get("value"), 1e3 * r.get("ms_per_launch", 0), r.get("frac", 0),
        1e3 * d["stages"]["cnn_ms_per_frame"], (r.get("cnn") or {}).get("launches_per_call")))
except Exception as e:
    print("${name}: no result", e)
PY
      ;;
    launches)
      args=$(echo "${a3:-}" | tr ',' ' ')
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --cache-control none -s 300 -c 400 --csv \
        --log-file gpurun_out/${TAG}_launches_${name}.csv python bench.py $args > gpurun_out/${TAG}_launches_${name}.log 2>&1 || tail -3 gpurun_out/${TAG}_launches_${name}.log ;;
    ncux)
      args=$(echo "${a6:-}" | tr ',' ' ')
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${a3}" -s ${a4} -c ${a5} -f -o gpurun_out/${TAG}_ncu_${name} \
        python bench.py $args > gpurun_out/${TAG}_ncu_${name}.log 2>&1 || tail -3 gpurun_out/${TAG}_ncu_${name}.log ;;
    ncu)
      args=$(echo "${a4:-}" | tr ',' ' ')
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:${a3} -s 6 -c 2 -f -o gpurun_out/${TAG}_ncu_${name} \
        python bench.py $args > gpurun_out/${TAG}_ncu_${name}.log 2>&1 || tail -3 gpurun_out/${TAG}_ncu_${name}.log ;;
  esac
done
