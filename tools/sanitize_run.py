"""Workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel of the library at
small sizes, checked against the oracle so that a sanitizer-clean run is also a correct one.
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
No torch import: only ctypes + numpy, so the report is about this library's kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dsvlibs as L

gpu, port = L.gpu(), L.port()
ref = L.ref() if L.have_ref() else None
rng = np.random.default_rng(0)
# kernel level: SBT forward / inverse, HZCC encode / decode, I and P, odd sub-band sizes
for (w, h) in ((176, 144), (134, 70)):
    pix = rng.integers(0, 256, size=(h, w + 2), dtype=np.uint8)
    nbh, nbv = (w + 15) // 16, (h + 15) // 16
    stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
    for isP in (0, 1):
        a = gpu.fwd_sbt(pix, w, h, w, h, isP)
        assert np.array_equal(a, port.fwd_sbt(pix, w, h, w, h, isP))
        co = (a // 9) * 9
        for c in (0, 1):
            assert np.array_equal(gpu.inv_sbt(co, 313, isP, c, w, h), port.inv_sbt(co, 313, isP, c, w, h))
        sa, ca = port.encode_plane(a, 313, isP, 0, stable, nbh, nbv)
        sb, cb = gpu.encode_plane(a, 313, isP, 0, stable, nbh, nbv)
        assert np.array_equal(sa, sb) and np.array_equal(ca, cb)
        assert np.array_equal(port.decode_plane(sa, w, h, 313, isP, 0, stable, nbh, nbv),
                              gpu.decode_plane(sa, w, h, 313, isP, 0, stable, nbh, nbv))
# motion: search + compensation, 4:2:0 and 4:4:4
for fmt in ("420", "444"):
    w, h = 176, 144
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 4, 0, start=3)
    fs = L.synth_sequence(w, h, fmt, 1, 4, 0, start=4)
    pe, me = gpu.hme(fs, fr, w, h, sub, 3)
    if ref:
        pr, mr = ref.hme(fs, fr, w, h, sub, 3)
        assert pr == pe and all(np.array_equal(mr[k], me[k]) for k in mr.dtype.names)
    mv = me.copy()
    mv["mode"] = (rng.random(mv.shape) < 0.3).astype(np.uint8)
    mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=mv.shape), 0).astype(np.uint8)
    pb, rb = gpu.sub_pred(mv, w, h, sub, fs, fr)
    if ref:
        pa, ra = ref.sub_pred(mv, w, h, sub, fs, fr)
        assert np.array_equal(pa, pb) and np.array_equal(ra, rb)
    gpu.add_pred(mv, w, h, sub, rb, fr)
# whole codec: per-picture API, batch API (2 lanes), chain-sharded long API, with a scene cut
w, h, fmt, n = 176, 144, "420", 8
sub = L.SUBSAMP[fmt]
fb = L.frame_bytes(w, h, sub)
yuv = L.synth_sequence(w, h, fmt, n, 1, 5)
cfg = L.make_cfg(w, h, fmt, gop=4)
stream, pk, _ = gpu.encode_sequence(cfg, yuv, n)
nf, dec, _, _ = gpu.decode_stream(stream, w, h, sub, n)
assert nf == n
if ref:
    assert ref.encode_sequence(cfg, yuv, n)[0] == stream
    assert np.array_equal(ref.decode_stream(stream, w, h, sub, n)[1], dec)
be = L.BatchEncoder(gpu, cfg, 2)
assert be.encode([yuv, yuv[:fb * 3]], 3) == [gpu.encode_sequence(cfg, yuv[:fb * 3], 3)[0]] * 2
assert be.encode_long(yuv, n)[0] == stream
be.close()
bd = L.BatchDecoder(gpu, 2)
out, fr = bd.decode_long(stream, fb, n)
assert fr == n and np.array_equal(out, dec)
for flags in ((1, 0), (0, 1)):  # overlay, 4:2:0 output conversion
    bd.set_draw_info(7 if flags[0] else 0)
    bd.set_out420p(flags[1])
    bd.decode([stream], fb, n)
bd.close()
print("sanitize_run ok")
