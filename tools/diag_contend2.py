"""Per-kernel device times of the device-resident encode with and without an unrelated bulk D2H stream."""
import ctypes as C, sys, time, threading
sys.path.insert(0, "tests")
import torch, dsvlibs as L
W, H, FMT, NFR, B = 1920, 1080, "420", 12, 64
gpu = L.gpu(); lib = gpu.lib
sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub); sb = fb * NFR
cfg = L.make_cfg(W, H, FMT, gop=12, qp=85)
d_yuv = torch.empty(B * sb, dtype=torch.uint8, device="cuda")
for s in range(B):
    lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * sb), 0)
cap = 8 << 20
h_str = torch.zeros(B * cap, dtype=torch.uint8).pin_memory()
enc = L.BatchEncoder(gpu, cfg, B, 0); enc.set_kernel_timing(1)
sp = [h_str.data_ptr() + s * cap for s in range(B)]
n = 1 << 28
hb = torch.empty(n, dtype=torch.uint8).pin_memory(); db = torch.empty(n, dtype=torch.uint8, device="cuda")
stop = [False]
def traffic(direction):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        while not stop[0]:
            if direction == "h2d": db.copy_(hb, non_blocking=True)
            else: hb.copy_(db, non_blocking=True)
            st.synchronize()
def do_enc():
    enc.encode_ptrs([d_yuv.data_ptr() + s * sb for s in range(B)], NFR, 1, sp, [cap] * B)
res = {}
for bg in (None, "d2h"):
    stop[0] = False
    th = threading.Thread(target=traffic, args=(bg,)) if bg else None
    if th: th.start(); time.sleep(0.2)
    do_enc(); torch.cuda.synchronize(); enc.kernel_times(reset=True); enc.stats(reset=True)
    t0 = time.perf_counter(); do_enc(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    res[bg] = (wall, enc.kernel_times(), enc.stats())
    stop[0] = True
    if th: th.join()
for bg in res:
    w, kt, st = res[bg]
    print("background %s: wall %.1f ms, sum of kernel times %.1f ms, host ms in steps %.1f" % (bg, w, sum(v["ms"] for v in kt.values()), st["host_ms"]))
for name in res[None][1]:
    a, b = res[None][1][name], res["d2h"][1].get(name, {"ms": 0, "launches": 1})
    print("  %-24s x%3d  %8.1f us -> %8.1f us per launch" % (name, a["launches"], 1e3 * a["ms"] / a["launches"], 1e3 * b["ms"] / max(b["launches"], 1)))
