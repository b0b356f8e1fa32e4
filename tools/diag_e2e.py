"""Where does the e2e time go: encode/decode x host/device arms, plus raw pinned PCIe bandwidth."""
import ctypes as C, sys, time
sys.path.insert(0, "tests")
import torch, dsvlibs as L
W, H, FMT, NFR, B = 1920, 1080, "420", 12, int(sys.argv[1]) if len(sys.argv) > 1 else 32
gpu = L.gpu(); lib = gpu.lib
sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub); sb = fb * NFR
cfg = L.make_cfg(W, H, FMT, gop=12, qp=85)
d_yuv = torch.empty(B * sb, dtype=torch.uint8, device="cuda")
for s in range(B):
    lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * sb), 0)
h_yuv = torch.empty(B * sb, dtype=torch.uint8).pin_memory(); h_yuv.copy_(d_yuv)
cap = 8 << 20
h_str = torch.zeros(B * cap, dtype=torch.uint8).pin_memory(); d_str = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty(B * sb, dtype=torch.uint8, device="cuda"); h_out = torch.empty(B * sb, dtype=torch.uint8).pin_memory()
enc = L.BatchEncoder(gpu, cfg, B, 0); dec = L.BatchDecoder(gpu, B, 0)
sp = [h_str.data_ptr() + s * cap for s in range(B)]; sdp = [d_str.data_ptr() + s * cap for s in range(B)]
def t(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, (rc, lens) = t(lambda: enc.encode_ptrs([h_yuv.data_ptr() + s * sb for s in range(B)], NFR, 0, sp, [cap] * B)); print("enc host   %.1f ms  %.0f pic/s" % (ms, B * NFR / ms * 1e3))
ms, _ = t(lambda: enc.encode_ptrs([d_yuv.data_ptr() + s * sb for s in range(B)], NFR, 1, sp, [cap] * B)); print("enc device %.1f ms  %.0f pic/s" % (ms, B * NFR / ms * 1e3))
d_str.copy_(h_str)
ms, _ = t(lambda: dec.decode_ptrs(sp, None, lens, [h_out.data_ptr() + s * sb for s in range(B)], [sb] * B, 0)); print("dec host   %.1f ms  %.0f pic/s" % (ms, B * NFR / ms * 1e3))
ms, _ = t(lambda: dec.decode_ptrs(sp, sdp, lens, [d_out.data_ptr() + s * sb for s in range(B)], [sb] * B, 1)); print("dec device %.1f ms  %.0f pic/s" % (ms, B * NFR / ms * 1e3))
ms, _ = t(lambda: d_yuv.copy_(h_yuv, non_blocking=True)); print("H2D %.1f ms %.1f GB/s" % (ms, B * sb / ms / 1e6))
ms, _ = t(lambda: h_out.copy_(d_out, non_blocking=True)); print("D2H %.1f ms %.1f GB/s" % (ms, B * sb / ms / 1e6))
enc.stats(reset=True); dec.stats(reset=True)
enc.encode_ptrs([d_yuv.data_ptr() + s * sb for s in range(B)], NFR, 1, sp, [cap] * B)
dec.decode_ptrs(sp, sdp, lens, [d_out.data_ptr() + s * sb for s in range(B)], [sb] * B, 1)
print("host ms inside steps: enc %.2f dec %.2f (per call)" % (enc.stats()["host_ms"], dec.stats()["host_ms"]))
