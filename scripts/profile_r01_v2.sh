# Round-1 profiles of the final kernels (run under gpurun).  Each ncu pass follows a plain run of the same command that
# exited 0.  13 of our kernels per step: k_step, k_compact, stem + 9 convs (k_conv8), k_head.  Every pass is bounded.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 1000 --no-cpu-baseline --no-graph"
timeout 120 $CMD > gpurun_out/plain_v2.log 2>&1 || { tail -5 gpurun_out/plain_v2.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 13300 -c 260 --csv --log-file gpurun_out/launches_v2.csv $CMD > gpurun_out/ncu_launches_v2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv8 -s 10020 -c 10 -o gpurun_out/prof_conv8 -f $CMD > gpurun_out/ncu_conv8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_head" -s 2010 -c 4 -o gpurun_out/prof_tree_v2 -f $CMD > gpurun_out/ncu_tree_v2.log 2>&1
tail -1 gpurun_out/plain_v2.log | cut -c1-150
python scripts/summarize_launches.py gpurun_out/launches_v2.csv | head -12
ls -la gpurun_out | tail -6
