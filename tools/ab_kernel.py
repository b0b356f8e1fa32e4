"""A/B of one kernel: runs the device-resident arm of the bench workload a few times and prints the engines' own
event timings.  DSV1_B200_LIB=<variant .so> python tools/ab_kernel.py [batch]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch, dsvlibs as L, bench
bench.GOP = int(os.environ.get("AB_GOP", bench.GOP))  # AB_GOP=0: intra-only pictures
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H, FMT, NFR = bench.W, bench.H, bench.FMT, bench.NFR
gpu = L.gpu(); lib = gpu.lib; sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub); seq = fb * NFR
cfg = L.make_cfg(W, H, FMT, gop=bench.GOP, qp=bench.QP)
d_yuv = torch.empty(B * seq, dtype=torch.uint8, device="cuda")
for s in range(B):
    lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * seq), 0)
cap = 8 << 20
h_streams = torch.zeros(B * cap, dtype=torch.uint8).pin_memory(); d_streams = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty(B * seq, dtype=torch.uint8, device="cuda")
enc = L.BatchEncoder(gpu, cfg, B, 0); dec = L.BatchDecoder(gpu, B, 0); enc.set_kernel_timing(1); dec.set_kernel_timing(1)
sp = [h_streams.data_ptr() + s * cap for s in range(B)]; sdp = [d_streams.data_ptr() + s * cap for s in range(B)]
import time
for it in range(4):
    if it == 1:
        enc.stats(reset=True); dec.stats(reset=True); enc.kernel_times(reset=True); dec.kernel_times(reset=True); torch.cuda.synchronize(); t0 = time.perf_counter()
    rc, lens = enc.encode_ptrs([d_yuv.data_ptr() + s * seq for s in range(B)], NFR, 1, sp, [cap] * B)
    if it == 0:
        d_streams.copy_(h_streams)
    rc, fr = dec.decode_ptrs(sp, sdp, lens, [d_out.data_ptr() + s * seq for s in range(B)], [seq] * B, 1)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
es, ds = enc.stats(), dec.stats()
import hashlib
print(os.path.basename(L.GPU_SO), "step %.2f ms" % (dt * 1e3),
      "bmc enc %.1f us dec %.1f us" % (1e3 * es["bmc_ms"] / max(es["bmc_launches"], 1), 1e3 * ds["bmc_ms"] / max(ds["bmc_launches"], 1)),
      "fwd %.1f inv %.1f/%.1f us" % (1e3 * es["sbt_fwd_ms"] / es["sbt_fwd_launches"], 1e3 * es["sbt_inv_ms"] / max(es["sbt_inv_launches"], 1), 1e3 * ds["sbt_inv_ms"] / ds["sbt_inv_launches"]),
      "out md5", hashlib.md5(d_out[:fb * 3].cpu().numpy().tobytes()).hexdigest()[:8])

ek, dk = enc.kernel_times(), dec.kernel_times()
want = sys.argv[2].split(",") if len(sys.argv) > 2 else []
for side, kt in (("enc", ek), ("dec", dk)):
    for name, v in kt.items():
        if not want or any(w in name for w in want):
            print("   %-26s %s  %8.1f us/launch  x%d" % (name, side, 1e3 * v["ms"] / v["launches"], v["launches"] / 3))
