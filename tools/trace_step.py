"""Activity-by-activity trace of a few encoder steps with host input (torch.profiler / CUPTI): start, duration, stream."""
import ctypes as C, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch, dsvlibs as L
from torch.profiler import profile, ProfilerActivity
W, H, FMT, NFR, B = 1920, 1080, "420", 12, 64
gpu = L.gpu(); lib = gpu.lib
sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub); sb = fb * NFR
cfg = L.make_cfg(W, H, FMT, gop=12, qp=85)
d_yuv = torch.empty(B * sb, dtype=torch.uint8, device="cuda")
for s in range(B):
    lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * sb), 0)
h_yuv = torch.empty(B * sb, dtype=torch.uint8).pin_memory(); h_yuv.copy_(d_yuv)
cap = 8 << 20
h_str = torch.zeros(B * cap, dtype=torch.uint8).pin_memory()
enc = L.BatchEncoder(gpu, cfg, B, 0)
sp = [h_str.data_ptr() + s * cap for s in range(B)]
def run(): return enc.encode_ptrs([h_yuv.data_ptr() + s * sb for s in range(B)], NFR, 0, sp, [cap] * B)
run(); run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
evs = []
for e in prof.events():
    tr = e.time_range
    dev = str(e.device_type).endswith("CUDA")
    evs.append((tr.start, tr.end - tr.start, "GPU" if dev else "cpu", e.name.replace("dsv::", "").split("(")[0][:40]))
evs.sort()
t0 = evs[0][0]
lo, hi = float(sys.argv[1]) if len(sys.argv) > 1 else 20000, float(sys.argv[2]) if len(sys.argv) > 2 else 30000
for s, d, where, name in evs:
    if lo <= s - t0 <= hi and (where == "GPU" or d > 20 or "ynchronize" in name):
        print("%9.0f us  %8.1f us  %s  %s" % (s - t0, d, where, name))
