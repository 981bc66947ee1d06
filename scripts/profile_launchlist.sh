mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 1000 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 18000 -c 420 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
wc -l gpurun_out/launches_final.csv
