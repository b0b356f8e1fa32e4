"""Device-resident encode / decode while an unrelated bulk copy saturates PCIe in one direction: how much of the host
arm's slowdown is PCIe contention on the control plane (zero-copy descriptor traffic, syncs)?"""
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
h_str = torch.zeros(B * cap, dtype=torch.uint8).pin_memory(); d_str = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty(B * sb, dtype=torch.uint8, device="cuda")
enc = L.BatchEncoder(gpu, cfg, B, 0); dec = L.BatchDecoder(gpu, B, 0)
sp = [h_str.data_ptr() + s * cap for s in range(B)]; sdp = [d_str.data_ptr() + s * cap for s in range(B)]
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
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
rc_lens = [None]
def do_enc():
    rc, lens = enc.encode_ptrs([d_yuv.data_ptr() + s * sb for s in range(B)], NFR, 1, sp, [cap] * B); rc_lens[0] = lens
do_enc(); d_str.copy_(h_str)
def do_dec():
    dec.decode_ptrs(sp, sdp, rc_lens[0], [d_out.data_ptr() + s * sb for s in range(B)], [sb] * B, 1)
for bg in (None, "h2d", "d2h"):
    stop[0] = False
    th = threading.Thread(target=traffic, args=(bg,)) if bg else None
    if th: th.start(); time.sleep(0.2)
    te, _ = t(do_enc); td, _ = t(do_dec)
    stop[0] = True
    if th: th.join()
    print("background %-5s: enc device %.1f ms, dec device %.1f ms" % (bg, te, td), flush=True)
# same, the lanes split over 2 / 4 engines on as many host threads
for parts in (2, 4):
    n2 = B // parts
    encs = [L.BatchEncoder(gpu, cfg, n2, 0) for _ in range(parts)]
    def part(i):
        encs[i].encode_ptrs([d_yuv.data_ptr() + s * sb for s in range(i * n2, (i + 1) * n2)], NFR, 1, sp[i * n2:(i + 1) * n2], [cap] * n2)
    def allp():
        th = [threading.Thread(target=part, args=(i,)) for i in range(parts)]
        [x.start() for x in th]; [x.join() for x in th]
    for bg in (None, "h2d", "d2h"):
        stop[0] = False
        th = threading.Thread(target=traffic, args=(bg,)) if bg else None
        if th: th.start(); time.sleep(0.2)
        te, _ = t(allp)
        stop[0] = True
        if th: th.join()
        print("%d engines x %d lanes, background %-5s: enc device %.1f ms" % (parts, n2, bg, te), flush=True)
    for e in encs: e.close()
