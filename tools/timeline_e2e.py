"""GPU timeline of the pipelined end-to-end arm (encode of step k+1 on one host thread while step k is decoded on
another, pinned host buffers): how busy are the SMs and the two copy directions, and where are they idle.
    python tools/timeline_e2e.py [--batch 64] [--steps 4] > profiles/rN_timeline_e2e.txt"""
import argparse, ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)


def union(iv):
    iv = sorted(iv)
    if not iv:
        return 0.0, []
    busy, gaps, cs, ce = 0.0, [], iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > ce:
            busy += ce - cs; gaps.append((ce, s)); cs, ce = s, e
        else:
            ce = max(ce, e)
    return busy + ce - cs, gaps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    import dsvlibs as L, bench
    W, H, FMT, NFR = bench.W, bench.H, bench.FMT, bench.NFR
    gpu = L.gpu(); lib = gpu.lib; B = args.batch
    sub = L.SUBSAMP[FMT]; fb = L.frame_bytes(W, H, sub); seq = fb * NFR
    cfg = L.make_cfg(W, H, FMT, gop=bench.GOP, qp=bench.QP)
    d_yuv = torch.empty(B * seq, dtype=torch.uint8, device="cuda")
    for s in range(B):
        lib.dsvb_synth_device(W, H, sub, 0, NFR, 100 + s, 0, C.c_void_p(d_yuv.data_ptr() + s * seq), 0)
    h_yuv = torch.empty(B * seq, dtype=torch.uint8).pin_memory(); h_yuv.copy_(d_yuv)
    cap = 8 << 20
    hs = [torch.zeros(B * cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_out = torch.empty(B * seq, dtype=torch.uint8).pin_memory()
    enc = L.BatchEncoder(gpu, cfg, B, 0); dec = L.BatchDecoder(gpu, B, 0)
    sp = [[h.data_ptr() + s * cap for s in range(B)] for h in hs]
    yp = [h_yuv.data_ptr() + s * seq for s in range(B)]; op = [h_out.data_ptr() + s * seq for s in range(B)]

    def run(steps):
        enc_done = [threading.Event() for _ in range(steps)]; dec_done = [threading.Event() for _ in range(steps)]
        lens = [None] * steps
        def ew():
            for k in range(steps):
                if k >= 2: dec_done[k - 2].wait()
                rc, lens[k] = enc.encode_ptrs(yp, NFR, 0, sp[k % 2], [cap] * B); enc_done[k].set()
        def dw():
            for k in range(steps):
                enc_done[k].wait()
                dec.decode_ptrs(sp[k % 2], None, lens[k], op, [seq] * B, 0); dec_done[k].set()
        a, b = threading.Thread(target=ew), threading.Thread(target=dw)
        a.start(); b.start(); a.join(); b.join()

    run(3); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter(); run(args.steps); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    k, h2d, d2h = [], [], []
    for e in prof.events():
        if str(e.device_type).endswith("CUDA"):
            tr = e.time_range; n = e.name.lower()
            (h2d if "htod" in n else d2h if "dtoh" in n else k).append((tr.start, tr.end))
    t_lo = min(x[0] for x in k + h2d + d2h); t_hi = max(x[1] for x in k + h2d + d2h)
    span = (t_hi - t_lo) / 1e3
    print("pipelined e2e, %d steps of %d x %d pictures: wall %.1f ms = %.1f ms per step (%.0f pictures/s); device span %.1f ms"
          % (args.steps, B, NFR, wall, wall / args.steps, B * NFR * args.steps / wall * 1e3, span))
    for name, iv in (("kernels (union)", k), ("H2D copies", h2d), ("D2H copies", d2h), ("anything", k + h2d + d2h)):
        busy, gaps = union(iv)
        big = sorted(((b - a) for a, b in gaps), reverse=True)
        print("  %-16s busy %7.1f ms = %5.1f %% of the span; %d gaps > 200 us totalling %.1f ms; largest %s us"
              % (name, busy / 1e3, 100 * busy / 1e3 / span, sum(1 for g in big if g > 200), sum(g for g in big if g > 200) / 1e3,
                 [round(g) for g in big[:5]]))
    print("  sum of kernel durations %.1f ms" % (sum(b - a for a, b in k) / 1e3))
    # 5 ms buckets: busy fraction of SMs / H2D / D2H
    nb = int(span / 5) + 1
    rows = []
    for name, iv in (("sm", k), ("h2d", h2d), ("d2h", d2h)):
        acc = [0.0] * nb
        for a, b in sorted(iv):
            a -= t_lo; b -= t_lo
            i = int(a / 5000)
            while a < b and i < nb:
                e = min(b, (i + 1) * 5000); acc[i] += e - a; a = e; i += 1
        rows.append((name, acc))
    print("  busy %% per 5 ms bucket (kernel durations summed over streams, can exceed 100):")
    for name, acc in rows:
        print("   %-4s %s" % (name, " ".join("%3d" % min(999, round(v / 50)) for v in acc)))


if __name__ == "__main__":
    main()
