"""profiles/r02_configs.json: the committed bench lines of every BASELINE config (verdict r01 item 5), collected from the
gpurun outputs of scripts/gpu_r03z.sh (1 GPU) and scripts/gpu_r03y.sh (8 GPUs), scripts/gpu_r03x.sh (2 GPUs); the 8-GPU training loop is
from scripts/gpu_r02k.sh)."""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(p):
    p = os.path.join(root, "gpurun_out", p)
    if not os.path.exists(p):
        return None
    for line in open(p):
        if line.startswith("{"):
            return json.loads(line)
out = {
    "note": "one bench.py JSON line per BASELINE.json config; clocks sampled during the timed region are inside every line",
    "configs[2] connect_four 800 sims 16384 trees, 1 GPU (python bench.py)": load("r03z_c4.json"),
    "configs[2] 8 GPUs (torchrun, --gpus 8 --steps 20 --warmup 5)": load("r03y_c4_n8.json"),
    "configs[2] 2 GPUs": load("r03x_c4_n2.json"),
    "configs[1] breakthrough 6x6 200 sims 1024 games shipped checkpoint, exact mode (--config bt6)": load("r03z_bt6.json"),
    "configs[1] virtual-loss mode K=8, NOT bit-exact (--config bt6 --virtual-loss 8)": load("r03z_bt6_vl8.json"),
    "configs[3] breakthrough 8x8 800 sims random-init net, 1 GPU (--config bt8)": load("r03z_bt8.json"),
    "configs[3] 8 GPUs (torchrun, --config bt8)": load("r03y_bt8_n8.json"),
    "configs[4] full train.py loop, 1 GPU (--config train)": load("r03z_train.json"),
    "configs[4] 8 GPUs (torchrun, --config train)": load("r02k_train_n8.json"),
    "reference arm (python bench.py --impl reference --gpus 1 --steps 20 --warmup 5): CPU port, all host cores": load("r03z_ref.json"),
    "configs[0] connect_four single game 100 sims on CPU": "cpu_baseline.single_process_100sims_sims_per_s of the configs[2] line",
}
json.dump(out, open(os.path.join(root, "profiles", "r02_configs.json"), "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict):
        print("%-95s %12.1f %s" % (k[:95], v["value"], v["unit"]))
