mkdir -p gpurun_out
set -x
python bench.py --steps 20 --warmup 330 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 500 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 330 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 320 -c 3 -o gpurun_out/prof_kstep_r01 python bench.py --steps 20 --warmup 330 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-300
ls -la gpurun_out
