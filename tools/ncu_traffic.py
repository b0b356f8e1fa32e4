"""profiles/traffic.json from `ncu --set full` captures: per kernel, dram__bytes_read.sum + dram__bytes_write.sum per launch
(mean over the captured launches) and the capture it came from.  bench.py reports it as roofline.traffic.
    python tools/ncu_traffic.py profiles/traffic.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, json, os, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out, reps = sys.argv[1], sys.argv[2:]
res = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("dsv::", "")
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]]) * UNIT[units[idx[m]]]
        dur = float(r[idx["gpu__time_duration.sum"]]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[units[idx["gpu__time_duration.sum"]]]
        per.setdefault(name, []).append((tot, dur))
    for name, v in per.items():
        res[name] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "ms_per_launch_under_ncu": sum(x[1] for x in v) / len(v),
                     "launches_captured": len(v), "source": os.path.basename(rep)}
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1, sort_keys=True))
