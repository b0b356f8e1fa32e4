import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # quiet under `| head`
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],2),"launches",d["gpu_launches"])
for k,v in d["kernels"].items(): print(" ",k, "ms/launch",round(v["ms_per_launch"],3), "GB/s",round(v["achieved_gbs"]), "frac",round(v["frac"],3), "share",round(v["share_of_step"],3))
if "cpu_baseline" in d and d["cpu_baseline"]: print(" cpu", d["cpu_baseline"]["value"])
