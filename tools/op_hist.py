import csv, subprocess, sys, re, collections
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
ie, src = h.index("Instructions Executed"), h.index("Source")
ops = collections.Counter(); tot = 0
for r in rows[hdr_i[0] + 1:end]:
    if len(r) > ie and r[ie].isdigit():
        s = r[src].strip()
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0].rstrip(";") if s else "?"
        base = ".".join(op.split(".")[:2])
        ops[base] += int(r[ie]); tot += int(r[ie])
print("total", tot)
for k, v in ops.most_common(40):
    print("%6.2f%%  %s" % (100.0 * v / tot, k))
