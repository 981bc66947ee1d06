#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r02t_train.json 2> gpurun_out/r02t_train.err
timeout 900 python bench.py --config train --generations 3 --eager-train > gpurun_out/r02t_train_eager.json 2> gpurun_out/r02t_train_eager.err
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r02t_train2.json 2> gpurun_out/r02t_train2.err
python - <<'PY'
import json
for f in ["train","train_eager","train2"]:
    for line in open("gpurun_out/r02t_%s.json" % f):
        if line.startswith("{"):
            l=json.loads(line); print(f, "gen %.3f train %.3f games/s %.1f" % (l["generation_s"], l["train_s"], l["value"]), l["train_step"])
PY
