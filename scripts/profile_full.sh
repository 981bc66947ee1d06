# usage: bash scripts/profile_full.sh <tag> <kernel-regex> <skip> [extra bench args]   (ncu --set full on one kernel)
mkdir -p gpurun_out
TAG=$1; KRE=$2; SKIP=$3; shift 3
python bench.py --steps 10 --warmup 330 --no-cpu-baseline --no-graph "$@" > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 3 -o gpurun_out/prof_$TAG -f python bench.py --steps 10 --warmup 330 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncufull_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-120
