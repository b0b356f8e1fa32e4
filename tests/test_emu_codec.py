"""Kernel-logic check in the GPU-less container: the SAME .cu / host sources compiled for the test-only CPU
emulator (tests/emu) drive a tiny sequence through dsv_enc / dsv_dec and the motion entry points, vs the
unmodified reference.  Small sizes only (one OS thread per CUDA thread); the parity tests proper are -m gpu."""
import subprocess

import numpy as np
import pytest

import dsvlibs as L


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", L.PKG, "emu"], check=True, stdout=subprocess.DEVNULL)
    return L.emu()


def test_codec_tiny(emu, ref):
    w, h, fmt, n = 96, 80, "420", 3
    yuv = L.synth_sequence(w, h, fmt, n, 2, 0)
    cfg = L.make_cfg(w, h, fmt, gop=12)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = emu.encode_sequence(cfg, yuv, n)
    assert pa == pb and sa == sb
    na, da, _, _ = ref.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    nb, db, _, _ = emu.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    assert na == nb == n and np.array_equal(da, db)


def test_motion_tiny(emu, ref):
    w, h, fmt = 112, 96, "444"
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 4, 0, start=3)
    fs = L.synth_sequence(w, h, fmt, 1, 4, 0, start=4)
    pr, mr = ref.hme(fs, fr, w, h, sub, 3)
    pe, me = emu.hme(fs, fr, w, h, sub, 3)
    assert pr == pe and all(np.array_equal(mr[k], me[k]) for k in mr.dtype.names)
    rng = np.random.default_rng(0)
    mv = mr.copy()
    mv["x"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
    mv["y"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
    mv["mode"] = (rng.random(mv.shape) < 0.3).astype(np.uint8)
    mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=mv.shape), 0).astype(np.uint8)
    pa, ra = ref.sub_pred(mv, w, h, sub, fs, fr)
    pb, rb = emu.sub_pred(mv, w, h, sub, fs, fr)
    assert np.array_equal(pa, pb) and np.array_equal(ra, rb)
    assert np.array_equal(ref.add_pred(mv, w, h, sub, ra, fr), emu.add_pred(mv, w, h, sub, ra, fr))


def test_hzcc_decode_small(emu, port):
    rng = np.random.default_rng(1)
    for cw, ch in [(16, 16), (136, 68)]:
        for isP in (0, 1):
            stable = rng.integers(0, 4, size=20, dtype=np.uint8)
            co = (rng.laplace(0, 400, size=(ch, cw)) * (rng.random((ch, cw)) < 0.2)).astype(np.int32)
            s, _ = port.encode_plane(co, 313, isP, 1, stable, 5, 4)
            assert np.array_equal(port.decode_plane(s, cw, ch, 313, isP, 1, stable, 5, 4),
                                  emu.decode_plane(s, cw, ch, 313, isP, 1, stable, 5, 4))


def test_decoder_output_options_tiny(emu, ref):
    """-drawinfo overlay (incl. intra dots that spill past the luma plane on edge blocks) and -out420p, through
    dsv_dec and through the batch decoder's device-side conversion, vs the reference CLI's procedure."""
    w, h, fmt, n = 72, 56, "444", 3
    sub = L.SUBSAMP[fmt]
    yuv = L.synth_sequence(w, h, fmt, n, 5, 2)
    cfg = L.make_cfg(w, h, fmt, gop=12, qp=60, do_scd=0, intra_pct=100)
    s, _, _ = ref.encode_sequence(cfg, yuv, n)
    plain = ref.decode_stream(s, w, h, sub, n)[1]
    for draw, to420 in ((7, 0), (5, 1)):
        na, da, _, _ = ref.decode_stream(s, w, h, sub, n, draw_info=draw, to_420p=to420)
        if not to420:
            nb, db, _, _ = emu.decode_stream(s, w, h, sub, n, draw_info=draw, to_420p=to420)
            assert na == nb == n and np.array_equal(da, db)
        bd = L.BatchDecoder(emu, 2)
        bd.set_draw_info(draw)
        bd.set_out420p(to420)
        outs, fr = bd.decode([s, s], L.frame_bytes(w, h, L.SUBSAMP["420"] if to420 else sub), n)
        bd.close()
        assert fr == [n, n] and np.array_equal(outs[0], da) and np.array_equal(outs[1], da)
        if draw and not to420:
            fb = L.frame_bytes(w, h, sub)
            a, p = da.reshape(n, fb), plain.reshape(n, fb)
            assert (a[:, :w * h] != p[:, :w * h]).any()
            assert (a[:, w * h:] != p[:, w * h:]).any()  # the reference's unchecked intra dots reached chroma


def test_batch_api_ragged_and_small_buffers(emu):
    """Batch API edge cases on the host side of the engines: streams of different lengths sharing lanes (a lane that
    runs out of packets idles while the others go on; more streams than lanes), an output buffer with room for only
    some pictures, a stream buffer that is too small."""
    w, h, fmt = 32, 32, "420"
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    cfg = L.make_cfg(w, h, fmt, gop=12, qp=70)
    counts = [2, 1, 1]
    seqs = [L.synth_sequence(w, h, fmt, n, 50 + i, 0) for i, n in enumerate(counts)]
    streams = [emu.encode_sequence(cfg, s, n)[0] for s, n in zip(seqs, counts)]
    want = [emu.decode_stream(s, w, h, sub, n)[1] for s, n in zip(streams, counts)]
    bd = L.BatchDecoder(emu, 2)
    outs, fr = bd.decode(streams, fb, 2)
    assert fr == counts
    for o, wnt, n in zip(outs, want, counts):
        assert np.array_equal(o[:fb * n], wnt)
    # room for the first picture only: later pictures are decoded (references stay in step) but not delivered
    bufs = [np.frombuffer(s, dtype=np.uint8) for s in streams]
    small = [np.zeros(fb, dtype=np.uint8) for _ in streams]
    rc, fr = bd.decode_ptrs([b.ctypes.data for b in bufs], None, [len(b) for b in bufs], [o.ctypes.data for o in small],
                            [fb] * 3, 0)
    bd.close()
    assert rc == 0 and fr == [1, 1, 1]
    for o, wnt in zip(small, want):
        assert np.array_equal(o, wnt[:fb])
    # encoder: a stream buffer that cannot hold the packets is reported, not overrun
    be = L.BatchEncoder(emu, cfg, 2)
    guard = np.full(4096, 0xAB, dtype=np.uint8)
    rc, lens = be.encode_ptrs([seqs[0].ctypes.data], 2, 0, [guard.ctypes.data], [64])
    be.close()
    assert rc == -1
    assert (guard[64:] == 0xAB).all()


def test_odd_geometry_tiny(emu, ref):
    """Widths that leave partial words / partial blocks at the right and bottom edges (block 16x16 on 54x38: the
    last block column is 6 samples wide): search, block statistics, reduced-range intra test, whole codec."""
    w, h, fmt, n = 54, 38, "420", 3
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 6, 0, start=0)
    fs = L.synth_sequence(w, h, fmt, 1, 8, 0, start=1).copy()
    y = fs[:w * h].reshape(h, w)
    y[:, w - 12:] = 255
    y[h - 9:, :] = 0
    pr, mr = ref.hme(fs, fr, w, h, sub, 3)
    pe, me = emu.hme(fs, fr, w, h, sub, 3)
    assert pr == pe and all(np.array_equal(mr[k], me[k]) for k in mr.dtype.names if k != "pad")
    assert (mr["mode"] == 1).any()
    yuv = L.synth_sequence(w, h, fmt, n, 6, 2)
    cfg = L.make_cfg(w, h, fmt, gop=12, qp=60, do_scd=0, intra_pct=100)
    sa, _, _ = ref.encode_sequence(cfg, yuv, n)
    sb, _, _ = emu.encode_sequence(cfg, yuv, n)
    assert sa == sb
    assert np.array_equal(ref.decode_stream(sa, w, h, sub, n)[1], emu.decode_stream(sa, w, h, sub, n)[1])


def test_motion_scene_cut_quadrants(emu, ref):
    """Source and reference from unrelated content (a scene cut): most blocks go through the cascade's intra branch,
    i.e. through the passes only intra blocks take (reduced-range test, quadrant good / evil metric with partial
    submasks), on 32-wide blocks (packed-word path) and with half-pel winners next to them."""
    w, h, fmt = 352, 288, "420"
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 4, 7, start=9)
    fs = L.synth_sequence(w, h, fmt, 1, 2, 0, start=4)
    pr, mr = ref.hme(fs, fr, w, h, sub, 3)
    pe, me = emu.hme(fs, fr, w, h, sub, 3)
    assert pr == pe and all(np.array_equal(mr[k], me[k]) for k in mr.dtype.names if k != "pad")
    assert (mr["mode"] == 1).sum() > 100 and ((mr["x"] | mr["y"]) & 1).any()
    assert len(set(mr["submask"].tolist()) - {0, 15}) > 0, "no partial quadrant mask in the sample"


def test_encoder_chunk_lists(emu, ref):
    """The encoder hands the HZCC passes a list scratch: dense chunks of the I picture leave per-lane lists, sparse
    chunks (the P pictures, the I picture's tail) one list in scan order, and the pack pass emits from them."""
    w, h, fmt, n = 176, 144, "420", 3
    yuv = L.synth_sequence(w, h, fmt, n, 0, 0)
    cfg = L.make_cfg(w, h, fmt, gop=12, qp=85)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = emu.encode_sequence(cfg, yuv, n)
    assert pa == pb and sa == sb


def test_dsv_hme_exported_interface_tiny(emu, ref):
    """The reference also exports dsv_hme(DSV_HME *): same call, caller-built pyramids, all levels compared."""
    w, h, fmt, lv = 90, 70, "444", 2
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 4, 0, start=3)
    fs = L.synth_sequence(w, h, fmt, 1, 9, 0, start=4)
    pr, mr = ref.hme_api(fs, fr, w, h, sub, lv)
    pe, me = emu.hme_api(fs, fr, w, h, sub, lv)
    assert pr == pe and all(np.array_equal(mr[k], me[k]) for k in mr.dtype.names if k != "pad")
    assert (mr[0]["mode"] == 1).any() and (mr[1]["x"] != 0).any()


def test_long_sequence_chain_sharding_tiny(emu, ref):
    """dsvb_encode_long / dsvb_decode_long: ONE sequence sharded by I-delimited chains over the lanes (phase A search
    of every picture, serial host pass, chains on lanes, ordered gather) is byte-identical to feeding the pictures
    to the reference one by one -- across GOP starts, a forced I picture at a scene cut (which does not restart the
    GOP) and a stability refresh; then the same over two (emulated) devices.  Larger cases: tests/test_gpu_long.py."""
    w, h, fmt, n = 32, 32, "420", 7
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, n, 7, 4)  # scene cut at picture 4
    cfg = L.make_cfg(w, h, fmt, gop=3, qp=70, stable_refresh=2)
    want, pk, _ = ref.encode_sequence(cfg, yuv, n)
    be = L.BatchEncoder(emu, cfg, 2)
    got, info = be.encode_long(yuv, n)
    be.close()
    assert got == want, info
    assert info[0] == 4 and info[1] == 1  # chains: GOP starts at 0, 3, 6 + the forced I picture at the cut
    _, wdec, _, _ = ref.decode_stream(want, w, h, sub, n)
    bd = L.BatchDecoder(emu, 2)
    out, fr = bd.decode_long(want, fb, n)
    bd.close()
    assert fr == n and np.array_equal(out, wdec)
    # two (emulated) devices: one engine + host thread per device, pictures / chains dealt in contiguous runs
    mg = L.MultiGpu(emu, cfg, 1, [0, 0])
    got, info = mg.encode_long(yuv[:fb * 5], 5)
    mg.close()
    assert got == ref.encode_sequence(cfg, yuv[:fb * 5], 5)[0] and info[3] == 2


def test_codec_interior_tiles(emu, ref):
    """384x192: the smallest picture with an interior 128x64 tile, i.e. the inverse transform's compile-time-geometry
    path, combined with the empty-band tile flags of P pictures (encoder: flags from the quantiser, reconstruction
    feeds the next P picture; decoder: flags from the scatter pass, flag-guided clean-up)."""
    w, h, fmt, n = 384, 192, "420", 3
    yuv = L.synth_sequence(w, h, fmt, n, 3, 0)
    cfg = L.make_cfg(w, h, fmt, gop=12)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = emu.encode_sequence(cfg, yuv, n)
    assert sa == sb
    na, da, _, _ = ref.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    nb, db, _, _ = emu.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    assert na == nb == n and np.array_equal(da, db)
