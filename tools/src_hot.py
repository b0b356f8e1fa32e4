"""Instruction / stall-sample share per CUDA source line of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):  python tools/src_hot.py rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = {}
fname, h, seen = None, None, 0
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        seen += 1
        continue
    if r and r[0] == "Line No":
        h = r
        ie, sm = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or seen > 1 and False:
        continue
    if len(r) > ie and r[0].isdigit() and r[ie].isdigit() and r[2] == "-":
        # CUDA line row (no SASS address): aggregated metrics of the line
        k = (fname, int(r[0]), r[1].strip())
        a = agg.setdefault(k, [0, 0])
        a[0] += int(r[ie])
        a[1] += int(r[sm]) if r[sm].isdigit() else 0
ti, ts = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print("warp instructions", ti, "stall samples", ts)
for (f, ln, txt), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% stall  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / ts, f, ln, txt[:110]))
