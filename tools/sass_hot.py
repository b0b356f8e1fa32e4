"""Hot SASS regions of one kernel in an .ncu-rep: python tools/sass_hot.py rep kernel_regex"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
ie, sm, src = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
data = [(int(r[ie]), int(r[sm]), r[src].strip()) for r in rows[hdr_i[0] + 1:end] if len(r) > ie and r[ie].isdigit()]
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("warp instructions", ti, "stall samples", ts, "sass lines", len(data))
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
for b in range(0, len(data), B):
    seg = data[b:b + B]
    i, s = sum(d[0] for d in seg), sum(d[1] for d in seg)
    if i == 0:
        continue
    ops = {}
    for n, _, t in seg:
        parts = t.split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + n
    top = " ".join("%s:%d%%" % (k, 100 * v / i) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    print("%5d-%5d instr %5.1f%% samples %5.1f%%  %s" % (b, b + B - 1, 100 * i / ti, 100 * s / ts, top))
print("top stall instructions:")
for d in sorted(data, key=lambda d: -d[1])[:12]:
    print("  %5d samples  exec %8d  %s" % (d[1], d[0], d[2][:100]))
