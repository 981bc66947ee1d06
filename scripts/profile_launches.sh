# usage: bash scripts/profile_launches.sh <tag> [extra bench args]
mkdir -p gpurun_out
TAG=$1; shift
python bench.py --steps 10 --warmup 330 --no-cpu-baseline --no-graph "$@" > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 10 --warmup 330 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
