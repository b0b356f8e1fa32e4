"""Real (non-serialised) GPU timeline of one bench step via torch.profiler (Kineto / CUPTI activity records):
per-kernel totals, GPU busy time vs wall, and the idle gaps between kernels, separately for the batch encoder and
the batch decoder.  ncu replays kernels one by one with cold caches; this shows the step as it actually runs.

    python tools/timeline.py [--batch 32] [--host] > profiles/rN_timeline.txt
"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--host", action="store_true", help="pinned host buffers in and out (the e2e arm)")
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    import dsvlibs as L
    import bench

    W, H, FMT, NFR, GOP, QP = bench.W, bench.H, bench.FMT, bench.NFR, bench.GOP, bench.QP
    gpu = L.gpu()
    lib = gpu.lib
    B = args.batch
    sub = L.SUBSAMP[FMT]
    fb = L.frame_bytes(W, H, sub)
    seq = fb * NFR
    cfg = L.make_cfg(W, H, FMT, gop=GOP, qp=QP)
    d_yuv = torch.empty(B * seq, dtype=torch.uint8, device="cuda")
    for s in range(B):
        lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * seq), 0)
    h_yuv = torch.empty(B * seq, dtype=torch.uint8).pin_memory()
    h_yuv.copy_(d_yuv)
    cap = 8 << 20
    h_streams = torch.zeros(B * cap, dtype=torch.uint8).pin_memory()
    d_streams = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(B * seq, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(B * seq, dtype=torch.uint8).pin_memory()
    enc = L.BatchEncoder(gpu, cfg, B, 0)
    dec = L.BatchDecoder(gpu, B, 0)
    sp = [h_streams.data_ptr() + s * cap for s in range(B)]
    sdp = [d_streams.data_ptr() + s * cap for s in range(B)]
    src = h_yuv if args.host else d_yuv
    dst = h_out if args.host else d_out

    def do_enc():
        rc, lens = enc.encode_ptrs([src.data_ptr() + s * seq for s in range(B)], NFR, 0 if args.host else 1, sp, [cap] * B)
        assert rc == 0
        return lens

    def do_dec(lens):
        rc, fr = dec.decode_ptrs(sp, None if args.host else sdp, lens, [dst.data_ptr() + s * seq for s in range(B)], [seq] * B,
                                 0 if args.host else 1)
        assert rc == 0 and all(f == NFR for f in fr)

    lens = do_enc()
    d_streams.copy_(h_streams)
    for _ in range(2):
        lens = do_enc()
        do_dec(lens)
    torch.cuda.synchronize()

    for name, fn in (("encode", do_enc), ("decode", lambda: do_dec(lens))):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
        evs = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA or str(e.device_type).endswith("CUDA"):
                tr = e.time_range
                evs.append((tr.start, tr.end, e.name))
        evs.sort()
        if not evs:
            print(name, "no device activity recorded")
            continue
        # union of busy intervals (kernels and copies may overlap across streams)
        busy, cur_s, cur_e = 0.0, evs[0][0], evs[0][1]
        gaps = []
        for s, e, _ in evs[1:]:
            if s > cur_e:
                busy += cur_e - cur_s
                gaps.append(s - cur_e)
                cur_s, cur_e = s, e
            else:
                cur_e = max(cur_e, e)
        busy += cur_e - cur_s
        span = evs[-1][1] - evs[0][0]
        kern = [x for x in evs if not x[2].lower().startswith("memcpy") and not x[2].lower().startswith("memset")]
        kbusy = sum(e - s for s, e, _ in kern)
        tot = {}
        for s, e, n in evs:
            n = n.split("(")[0].replace("dsv::", "")
            t = tot.setdefault(n, [0, 0.0])
            t[0] += 1
            t[1] += e - s
        print("== %s of %d x %d pictures (%s buffers): wall %.2f ms, device span %.2f ms, busy (union) %.2f ms = %.0f %%, "
              "kernel time (sum) %.2f ms, %d activities" % (name, B, NFR, "host" if args.host else "device", wall, span / 1e3,
                                                            busy / 1e3, 100.0 * busy / span, kbusy / 1e3, len(evs)))
        gaps.sort(reverse=True)
        print("   idle: total %.2f ms in %d gaps; > 100 us: %d (%.2f ms); largest %s us" % (
            sum(gaps) / 1e3, len(gaps), sum(1 for g in gaps if g > 100), sum(g for g in gaps if g > 100) / 1e3,
            [round(g) for g in gaps[:6]]))
        for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
            print("   %-34s %5d x %8.1f us = %8.2f ms  %5.1f %%" % (n[:34], c, t / c, t / 1e3, 100.0 * t / span))


if __name__ == "__main__":
    main()
