"""Additive batch API (dsvb_*, include/dsv1_b200_batch.h): lanes in lock step produce exactly the bytes the
per-picture API produces; device-resident inputs/outputs; the CUDA synthetic-content generator equals
oracle/synth.c."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")))


def test_synth_device_matches_oracle(gpu):
    import torch
    for (w, h, fmt, n, seed, cut, start) in [(352, 288, "420", 3, 1, 2, 0), (176, 144, "444", 2, 4, 0, 13),
                                             (176, 144, "411", 2, 6, 0, 1), (1920, 1080, "420", 1, 2, 0, 5)]:
        sub = L.SUBSAMP[fmt]
        fb = L.frame_bytes(w, h, sub)
        d = torch.zeros(fb * n, dtype=torch.uint8, device="cuda")
        gpu.lib.dsvb_synth_device(w, h, sub, start, n, seed, cut, C.c_void_p(d.data_ptr()), 0)
        assert np.array_equal(d.cpu().numpy(), L.synth_sequence(w, h, fmt, n, seed, cut, start=start))


@pytest.mark.parametrize("case", [(352, 288, "420", 13, 12, 5, 3), (176, 144, "444", 9, 4, 7, 4), (640, 360, "420", 6, 0, 3, 2)])
def test_batch_equals_per_picture_api(gpu, case):
    w, h, fmt, n, gop, nseq, lanes = case
    sub = L.SUBSAMP[fmt]
    cfg = L.make_cfg(w, h, fmt, gop=gop)
    seqs = [L.synth_sequence(w, h, fmt, n, 40 + s, 7 if s == 1 else 0) for s in range(nseq)]
    want = [gpu.encode_sequence(cfg, s, n)[0] for s in seqs]
    be = L.BatchEncoder(gpu, cfg, lanes)
    got = be.encode(seqs, n)
    be.close()
    assert got == want
    bd = L.BatchDecoder(gpu, lanes)
    outs, fr = bd.decode(want, L.frame_bytes(w, h, sub), n)
    bd.close()
    assert fr == [n] * nseq
    for s, o in zip(want, outs):
        nf, d, _, _ = gpu.decode_stream(s, w, h, sub, n)
        assert nf == n and np.array_equal(d, o)


def test_batch_hd_golden_device_resident(gpu):
    """HD gop12 golden vector through the batch API with device-resident pictures in and out."""
    import torch
    g = GOLD["hd_gop12"]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    d_yuv = torch.zeros(fb * n, dtype=torch.uint8, device="cuda")
    gpu.lib.dsvb_synth_device(w, h, sub, 0, n, g["seed"], g["cut"], C.c_void_p(d_yuv.data_ptr()), 0)
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"])
    be = L.BatchEncoder(gpu, cfg, 2)
    out = np.zeros(2 * (16 << 20), dtype=np.uint8)
    ptrs = [out.ctypes.data, out.ctypes.data + (16 << 20)]
    rc, lens = be.encode_ptrs([d_yuv.data_ptr(), d_yuv.data_ptr()], n, 1, ptrs, [16 << 20] * 2)
    be.close()
    assert rc == 0 and lens[0] == lens[1] == g["dsv_len"]
    s0 = out[:lens[0]].tobytes()
    assert hashlib.md5(s0).hexdigest() == g["dsv_md5"]
    assert out[16 << 20:(16 << 20) + lens[1]].tobytes() == s0
    d_stream = torch.from_numpy(out[:lens[0]].copy()).cuda()
    d_out = torch.zeros(fb * n, dtype=torch.uint8, device="cuda")
    bd = L.BatchDecoder(gpu, 2)
    rc, fr = bd.decode_ptrs([ptrs[0]], [d_stream.data_ptr()], [lens[0]], [d_out.data_ptr()], [fb * n], 1)
    st = bd.stats()
    bd.close()
    assert rc == 0 and fr == [n]
    assert hashlib.md5(d_out.cpu().numpy().tobytes()).hexdigest() == g["dec_md5"]
    assert st["h2d_bytes"] == 0 and st["d2h_bytes"] == 0   # nothing crossed PCIe for the pictures or the packets


def test_batch_contiguous_host_buffers(gpu):
    """Sequences back to back in ONE host buffer (constant distance): the single strided-copy ingest / egress path."""
    w, h, fmt, n, nseq, lanes = 352, 288, "420", 7, 6, 4
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    cfg = L.make_cfg(w, h, fmt, gop=12)
    big = np.concatenate([L.synth_sequence(w, h, fmt, n, 60 + s, 0) for s in range(nseq)])
    want = [gpu.encode_sequence(cfg, big[s * n * fb:(s + 1) * n * fb], n)[0] for s in range(nseq)]
    cap = 1 << 20
    outs = np.zeros(nseq * cap, dtype=np.uint8)
    be = L.BatchEncoder(gpu, cfg, lanes)
    rc, lens = be.encode_ptrs([big.ctypes.data + s * n * fb for s in range(nseq)], n, 0,
                              [outs.ctypes.data + s * cap for s in range(nseq)], [cap] * nseq)
    be.close()
    assert rc == 0
    got = [outs[s * cap:s * cap + lens[s]].tobytes() for s in range(nseq)]
    assert got == want
    dec_all = np.zeros(nseq * n * fb, dtype=np.uint8)
    bd = L.BatchDecoder(gpu, lanes)
    rc, fr = bd.decode_ptrs([outs.ctypes.data + s * cap for s in range(nseq)], None, lens,
                            [dec_all.ctypes.data + s * n * fb for s in range(nseq)], [n * fb] * nseq, 0)
    bd.close()
    assert rc == 0 and fr == [n] * nseq
    for s in range(nseq):
        nf, d, _, _ = gpu.decode_stream(want[s], w, h, sub, n)
        assert np.array_equal(d, dec_all[s * n * fb:(s + 1) * n * fb])


def test_tile_flags_follow_the_quantiser(gpu):
    """Tile band flags (sbt.cuh): one byte per 128x64 tile, bit 0 = the tile's level-1 band blocks hold a non-zero
    coefficient, bit 1 = its level-2 blocks do.  At qp85 the level-1 quantiser step of a P picture (2^10, 2^9 in stable
    blocks) exceeds most level-1 coefficients of a residual (|LH| <= 4 * 255): the flags say so and the inverse
    transform, the HZCC scan and the decoder's clean-up skip those blocks; an I picture is dense at every level.  The
    streams are the reference's either way (checked by every other test in this file)."""
    w, h, fmt = 640, 384, "420"
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, 3, 21, 0)
    ntiles = (w // 128) * (h // 64) + 2 * ((w // 2 + 127) // 128) * ((h // 2 + 63) // 64)
    buf = np.zeros(4096, dtype=np.uint8)

    def flags_after(qp, nframes):
        be = L.BatchEncoder(gpu, L.make_cfg(w, h, fmt, gop=12, qp=qp), 1)
        be.encode([yuv[:fb * nframes]], nframes)
        n = gpu.lib.dsvb_enc_tile_flags(be.h, 0, C.c_void_p(buf.ctypes.data), len(buf))
        be.close()
        assert n == ntiles
        return buf[:n].copy()

    f_i = flags_after(85, 1)
    assert (f_i & 1).all() and (f_i & 2).all()          # I picture: every level-1 block flagged, level 2 dense as well
    f_p = flags_after(85, 3)
    luma = f_p[:(w // 128) * (h // 64)]
    assert (luma & 1).sum() <= len(luma) // 4            # P picture at qp85: (nearly) no level-1 block holds anything
    f_hq = flags_after(100, 3)
    assert (f_hq & 1).sum() > (f_p & 1).sum()            # a finer quantiser keeps level-1 coefficients alive
