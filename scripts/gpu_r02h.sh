#!/bin/bash
# round 2, pass h: device trainer (graph step), train config, ncu capture of one evaluation (traffic), launch list of a bench run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02h_pytest.log
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r02h_bench_train.json 2> gpurun_out/r02h_bench_train.err
timeout 900 python bench.py --config train --generations 3 --eager-train > gpurun_out/r02h_bench_train_eager.json 2> gpurun_out/r02h_bench_train_eager.err
python scripts/eval_profile_target.py > gpurun_out/r02h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_conv8|k_head' -s 11 -c 11 -f -o gpurun_out/r02h_eval python scripts/eval_profile_target.py > gpurun_out/r02h_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r02h_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 260 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r02h_ncu2.log 2>&1
tail -3 gpurun_out/r02h_pytest.log
