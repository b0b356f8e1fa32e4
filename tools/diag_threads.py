"""Does splitting a GPU's lanes over several engines / host threads fill the sync gaps?  python tools/diag_threads.py"""
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
h_yuv = torch.empty(B * sb, dtype=torch.uint8).pin_memory(); h_yuv.copy_(d_yuv)
cap = 8 << 20
h_str = [torch.zeros(B * cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
h_out = torch.empty(B * sb, dtype=torch.uint8).pin_memory()
for parts in (1, 2, 4):
    n = B // parts
    encs = [L.BatchEncoder(gpu, cfg, n, 0) for _ in range(parts)]
    decs = [L.BatchDecoder(gpu, n, 0) for _ in range(parts)]
    lens = [[None] * parts for _ in range(2)]
    def enc_part(i, k):
        rc, ln = encs[i].encode_ptrs([h_yuv.data_ptr() + s * sb for s in range(i * n, (i + 1) * n)], NFR, 0,
                                     [h_str[k % 2].data_ptr() + s * cap for s in range(i * n, (i + 1) * n)], [cap] * n)
        assert rc == 0; lens[k % 2][i] = ln
    def dec_part(i, k):
        rc, fr = decs[i].decode_ptrs([h_str[k % 2].data_ptr() + s * cap for s in range(i * n, (i + 1) * n)], None, lens[k % 2][i],
                                     [h_out.data_ptr() + s * sb for s in range(i * n, (i + 1) * n)], [sb] * n, 0)
        assert rc == 0
    def par(f, k):
        th = [threading.Thread(target=f, args=(i, k)) for i in range(parts)]
        [t.start() for t in th]; [t.join() for t in th]
    def timeit(f, reps=3):
        f(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): f()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
    te = timeit(lambda: par(enc_part, 0)); td = timeit(lambda: par(dec_part, 0))
    # pipelined: encode k+1 (all parts) while decode k (all parts)
    def pipe(steps=4):
        par(enc_part, 0)
        for k in range(steps):
            a = threading.Thread(target=par, args=(dec_part, k)); b = threading.Thread(target=par, args=(enc_part, k + 1))
            a.start(); b.start(); a.join(); b.join()
    tp = timeit(lambda: pipe(4), reps=2) / 5
    print("%d engine(s) x %d lanes per direction: enc host %.1f ms, dec host %.1f ms, pipelined %.1f ms per step (%.0f pictures/s)"
          % (parts, n, te, td, tp, B * NFR / tp * 1e3), flush=True)
    for e in encs: e.close()
    for d in decs: d.close()
