"""profiles/conv_traffic.json from an `ncu --set full` capture of one evaluation (scripts/eval_profile_target.py):
dram__bytes_read.sum + dram__bytes_write.sum per conv launch, averaged over the conv launches (k_conv8<false,...>), and the
per-launch table as profiles/<name>_raw.csv.  Usage: summarize_ncu_traffic.py <rep.ncu-rep> <game> <boards> <name>"""
import csv, io, json, os, subprocess, sys
rep, game, boards, name = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
out_rows, conv = [], []
for r in data:
    out_rows.append([r[ix[k]] for k in keep])
    nm = r[ix["Kernel Name"]]
    b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
        to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    if "k_conv8<0" in nm.replace("(bool)", "") or "k_conv8<false" in nm:
        conv.append(b)
with open(os.path.join(root, "profiles", name + "_raw.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[ix[k]] for k in keep])
    w.writerows(out_rows)
path = os.path.join(root, "profiles", "conv_traffic.json")
tab = json.load(open(path)) if os.path.exists(path) else {}
tab["%s/%d" % (game, boards)] = {"bytes_per_launch": sum(conv) / len(conv), "conv_launches": len(conv),
                                 "source": "profiles/%s_raw.csv (ncu --set full, one evaluation)" % name}
json.dump(tab, open(path, "w"), indent=1)
print(json.dumps(tab["%s/%d" % (game, boards)]))
