#!/bin/bash
# r03l: ncu launch list of the default benchmark step with the final r02 code (same command as scripts/gpu_r02h.sh)
mkdir -p gpurun_out; rm -f gpurun_out/r03l_*
timeout 300 python bench.py --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r03l_plain.json 2> gpurun_out/r03l_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 260 --csv --log-file gpurun_out/r03l_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r03l_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r03l_launches.csv
