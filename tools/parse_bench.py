"""one bench.py JSON line (stdin) -> a short table"""
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # quiet under `| head`
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.1f  e2e %.1f  ms/step %.2f (e2e %.2f, pcie floor %s)  enc %s dec %s  launches %s" % (
    d["value"], e["value"], d["ms_per_step"], e["ms_per_step"], e.get("pcie_floor_ms"), d.get("encode_pictures_per_s"),
    d.get("decode_pictures_per_s"), d["gpu_launches"]))
r = d.get("roofline") or {}
print("roofline", r.get("kernel"), r.get("frac"), r.get("north_star_kernels"))
for k, v in d["kernels"].items():
    print("  %-28s %8.1f us/launch x%4d  share %.3f  frac %s" % (k, v["ms_per_launch"] * 1e3, v["launches"], v["share_of_step"],
                                                                   ("%.3f" % v["frac"]) if v.get("frac") else "-"))
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["value"], d.get("verified", {}).get("what", "")[:60])
