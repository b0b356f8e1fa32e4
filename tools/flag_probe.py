"""GPU probe: per-kernel times of P pictures only = (I, P, P) run minus (I) run, 64 lanes of bench content"""
import ctypes as C, sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch, dsvlibs as L
gpu = L.gpu(); lib = gpu.lib
W, H, FMT = 1920, 1080, "420"
sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub)
B, NFR = 64, 3
cfg = L.make_cfg(W, H, FMT, gop=12, qp=85)
d = torch.empty(B * NFR * fb, dtype=torch.uint8, device="cuda")
for s in range(B):
    lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d.data_ptr() + s * NFR * fb), 0)
enc = L.BatchEncoder(gpu, cfg, B, 0); dec = L.BatchDecoder(gpu, B, 0); enc.set_kernel_timing(1); dec.set_kernel_timing(1)
cap = 8 << 20
out = torch.zeros(B * cap, dtype=torch.uint8).pin_memory()
d_out = torch.empty(B * NFR * fb, dtype=torch.uint8, device="cuda")
sp = [out.data_ptr() + s * cap for s in range(B)]
res = {}
for nfr in (1, 3, 1, 3):
    enc.kernel_times(reset=True); dec.kernel_times(reset=True)
    rc, lens = enc.encode_ptrs([d.data_ptr() + s * NFR * fb for s in range(B)], nfr, 1, sp, [cap] * B)
    rc, fr = dec.decode_ptrs(sp, None, lens, [d_out.data_ptr() + s * NFR * fb for s in range(B)], [NFR * fb] * B, 1)
    torch.cuda.synchronize()
    res[nfr] = (enc.kernel_times(), dec.kernel_times())
for side in (0, 1):
    for name in res[3][side]:
        a = res[3][side][name]["ms"]; b = res[1][side].get(name, {"ms": 0})["ms"]
        print("%-24s %s  I %8.1f us   P %8.1f us" % (name, "enc" if side == 0 else "dec", 1e3 * b, 1e3 * (a - b) / 2))
