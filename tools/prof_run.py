"""Tiny driver for ncu captures: python tools/prof_run.py <w> <h> <fmt> <frames> <gop> [enc|dec|both]"""
import sys
sys.path.insert(0, "tests")
import dsvlibs as L
w, h, fmt, n, gop = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
what = sys.argv[6] if len(sys.argv) > 6 else "both"
gpu = L.gpu()
yuv = L.synth_sequence(w, h, fmt, n, 2, 0)
cfg = L.make_cfg(w, h, fmt, gop=gop)
if what in ("enc", "both"):
    s, pk, sec = gpu.encode_sequence(cfg, yuv, n)
    print("enc fps", n / sec)
else:
    s, pk, sec = L.ref().encode_sequence(cfg, yuv, n)
if what in ("dec", "both"):
    nf, dec, meta, dsec = gpu.decode_stream(s, w, h, L.SUBSAMP[fmt], n)
    print("dec fps", n / dsec)
