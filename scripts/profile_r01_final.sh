# Round-1 final profiles (run under gpurun).  Each ncu pass follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 1000 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20300 -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv -s 10030 -c 3 -o gpurun_out/prof_conv_final -f $CMD > gpurun_out/ncu_conv_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_step|k_compact" -s 2010 -c 4 -o gpurun_out/prof_tree_final -f $CMD > gpurun_out/ncu_tree_final.log 2>&1
tail -1 gpurun_out/plain_final.log | cut -c1-150
ls -la gpurun_out | tail -8
