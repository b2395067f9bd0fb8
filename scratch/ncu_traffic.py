"""profiles/ncu_traffic.json from an `ncu --set full` capture of `bench.py --skip-extras` (raw CSV page):
    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > raw.csv ; python scratch/ncu_traffic.py raw.csv <rows> <source label>
bench.py reads the file for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch).
    python scratch/ncu_traffic.py raw.csv <rows> <source label> --sorted <steps in the capture>
adds the entry "cox_sorted_step": every kernel of the SORTED path (scratch/sorted_ncu.py) summed per fwd+bwd step; the
other entries are kept."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = "cox_binned_fwd_fused" if "cox_binned_fwd_fused" in name else ("cox_binned_bwd" if "cox_binned_bwd" in name else None)
    if key is None:
        continue
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    dur = float(r[idx["gpu__time_duration.sum"]])
    out.setdefault(key, []).append((rd, wr, dur))
res = {}
for k, v in out.items():
    n = len(v)
    res[k] = {"dram_bytes_read": sum(x[0] for x in v) / n, "dram_bytes_write": sum(x[1] for x in v) / n,
              "gpu_time_us": sum(x[2] for x in v) / n, "launches_averaged": n, "rows": int(sys.argv[2]), "source": sys.argv[3]}
if "--sorted" in sys.argv:
    steps = int(sys.argv[sys.argv.index("--sorted") + 1])
    tot = [0.0, 0.0, 0.0]
    per = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if "b200surv" not in name and "sortscan" not in name:
            continue
        rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        dur = float(r[idx["gpu__time_duration.sum"]])
        tot[0] += rd; tot[1] += wr; tot[2] += dur
        k = name.split("(")[0].split("::")[-1]
        per[k] = per.get(k, 0.0) + (rd + wr) / steps
    res = json.load(open("profiles/ncu_traffic.json"))
    res["cox_sorted_step"] = {"dram_bytes_read": tot[0] / steps, "dram_bytes_write": tot[1] / steps, "gpu_time_us": tot[2] / steps,
                              "launches_averaged": steps, "rows": int(sys.argv[2]), "source": sys.argv[3],
                              "per_kernel_bytes": {k: round(v) for k, v in per.items()}}
json.dump(res, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(res, indent=1))
