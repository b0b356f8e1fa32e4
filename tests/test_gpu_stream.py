"""Whole codec through the drop-in API (dsv_enc / dsv_dec, driven by tools/api_harness.c exactly like the
reference CLI drives them) on the B200 vs the committed golden vectors of the UNMODIFIED reference
(tests/golden/streams.json, made by tests/golden/make_golden.py): .dsv byte-for-byte (md5 + length), decoded
YUV exact (md5), plus cross-decoding against the reference library when it travelled with the snapshot."""
import hashlib
import json
import os

import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_golden_stream(gpu, name):
    g = GOLD[name]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], g["cut"])
    assert hashlib.md5(yuv.tobytes()).hexdigest() == g["yuv_md5"]
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"])
    stream, pk, _ = gpu.encode_sequence(cfg, yuv, n)
    assert len(stream) == g["dsv_len"] and len(pk) == g["packets"]
    assert hashlib.md5(stream).hexdigest() == g["dsv_md5"]
    # the stream is now known to equal the reference's: decode it on the GPU
    nf, dec, meta, _ = gpu.decode_stream(stream, w, h, L.SUBSAMP[fmt], n)
    assert nf == n and meta[:3] == [w, h, L.SUBSAMP[fmt]]
    assert hashlib.md5(dec.tobytes()).hexdigest() == g["dec_md5"]


@pytest.mark.parametrize("case", [(352, 288, "420", 20, 21, 9, 12, 30), (640, 360, "420", 10, 22, 0, 5, 100),
                                  (176, 144, "444", 12, 23, 6, 12, 0), (960, 540, "422", 6, 24, 0, 3, 85)])
def test_vs_reference_live(gpu, ref, case):
    """Content / parameter combinations outside the golden set, checked against the reference library."""
    w, h, fmt, n, seed, cut, gop, qp = case
    yuv = L.synth_sequence(w, h, fmt, n, seed, cut)
    cfg = L.make_cfg(w, h, fmt, gop=gop, qp=qp)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = gpu.encode_sequence(cfg, yuv, n)
    assert pa == pb
    assert sa == sb
    na, da, _, _ = ref.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    nb, db, _, _ = gpu.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    assert na == nb == n
    assert np.array_equal(da, db)


def test_closed_loop(gpu):
    """Encoder reconstruction == decoder output is implied by P frames decoding exactly; here: a noisy
    high-motion clip where every P frame leans on the previous reconstruction."""
    w, h, fmt, n = 352, 288, "420", 16
    yuv = L.synth_sequence(w, h, fmt, n, 31, 0)
    cfg = L.make_cfg(w, h, fmt, gop=15, qp=60)
    stream, _, _ = gpu.encode_sequence(cfg, yuv, n)
    nf, dec, _, _ = gpu.decode_stream(stream, w, h, L.SUBSAMP[fmt], n)
    assert nf == n
    err = np.abs(dec.astype(np.int32) - yuv.astype(np.int32)).mean()
    assert err < 12.0   # drift would blow this up


def test_decoder_error_paths(gpu):
    """dsv_decoder.c:300-331: bad FourCC -> error (frame count unchanged); pictures before metadata are skipped."""
    w, h, fmt, n = 176, 144, "420", 2
    yuv = L.synth_sequence(w, h, fmt, n, 1, 0)
    stream, pk, _ = gpu.encode_sequence(L.make_cfg(w, h, fmt, gop=0), yuv, n)
    # drop the first metadata packet: the first picture has no metadata -> skipped, second GOP decodes
    nf, dec, meta, _ = gpu.decode_stream(stream[pk[0]:], w, h, L.SUBSAMP[fmt], n)
    assert nf == 1
    bad = bytearray(stream)
    bad[pk[0]] = ord("X")   # corrupt the FourCC of the first picture packet
    nf, _, _, _ = gpu.decode_stream(bytes(bad), w, h, L.SUBSAMP[fmt], n)
    assert nf <= 1


@pytest.mark.parametrize("case", [(352, 288, "420", 40, 26, 17, 12, 800000), (640, 360, "420", 20, 27, 0, 12, 2500000)])
def test_abr_rate_control_vs_reference(gpu, ref, case):
    """ABR (the CLI's default mode, dsv_encoder.c:84-160,816-848): every picture's quantiser depends on the sizes of the
    packets before it, so a single byte of difference anywhere would snowball -- the stream must still be identical."""
    w, h, fmt, n, seed, cut, gop, bitrate = case
    yuv = L.synth_sequence(w, h, fmt, n, seed, cut)
    cfg = L.make_cfg(w, h, fmt, gop=gop, qp=70, rc_mode=1, bitrate=bitrate, quality=L.qp_to_quality(70) * 3 // 2)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = gpu.encode_sequence(cfg, yuv, n)
    assert pa == pb and sa == sb
    na, da, _, _ = ref.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    nb, db, _, _ = gpu.decode_stream(sa, w, h, L.SUBSAMP[fmt], n)
    assert na == nb == n and np.array_equal(da, db)


def test_corrupt_streams_do_not_hang_or_crash(gpu):
    """Bit flips / truncation inside picture packets: the decoder must come back (any return code) in bounded time."""
    w, h, fmt, n = 352, 288, "420", 6
    yuv = L.synth_sequence(w, h, fmt, n, 33, 0)
    stream, pk, _ = gpu.encode_sequence(L.make_cfg(w, h, fmt, gop=12), yuv, n)
    rng = np.random.default_rng(5)
    off = [0]
    for p in pk:
        off.append(off[-1] + p)
    for trial in range(12):
        bad = bytearray(stream)
        k = int(rng.integers(1, len(pk) - 1))            # some packet after the first metadata packet
        lo, hi = off[k] + 14, off[k + 1]
        if hi - lo < 8:
            continue
        for _ in range(int(rng.integers(1, 40))):
            at = int(rng.integers(lo, hi))
            bad[at] ^= 1 << int(rng.integers(0, 8))
        nf, _, _, _ = gpu.decode_stream(bytes(bad), w, h, L.SUBSAMP[fmt], n)
        assert -8 <= nf <= n
    # zeroed coefficient area: never-ending exp-Golomb prefixes
    bad = bytearray(stream)
    bad[off[1] + 64:off[2]] = bytes(off[2] - off[1] - 64)
    nf, _, _, _ = gpu.decode_stream(bytes(bad), w, h, L.SUBSAMP[fmt], n)
    assert nf <= n


OUTG = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "output_options.json")))


@pytest.mark.parametrize("name", sorted(OUTG))
def test_decoder_output_options_golden(gpu, name):
    """DSV_DECODER.draw_info (CLI -drawinfo; dsv_decoder.c:147-243) painted by overlay.cu, and the CLI's -out420p
    conversion, vs md5s of the unmodified reference (tests/golden/output_options.json)."""
    g = OUTG[name]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    sub = L.SUBSAMP[fmt]
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], g["cut"])
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"], **g["cfg"])
    stream, _, _ = gpu.encode_sequence(cfg, yuv, n)
    assert hashlib.md5(stream).hexdigest() == g["dsv_md5"]
    for key, md5 in sorted(g["dec"].items()):
        draw, to420 = int(key[4]), int(key[-1])
        nf, dec, _, _ = gpu.decode_stream(stream, w, h, sub, n, draw_info=draw, to_420p=to420)
        assert nf == n and hashlib.md5(dec.tobytes()).hexdigest() == md5, key
    # batch decoder: overlay and conversion both on the device, host-packed and device-resident destinations
    import torch
    for draw, to420 in ((7, 1), (7, 0), (0, 1)):
        want = g["dec"]["draw%d_420p%d" % (draw, to420)]
        fb = L.frame_bytes(w, h, L.SUBSAMP["420"] if to420 else sub)
        bd = L.BatchDecoder(gpu, 2)
        bd.set_draw_info(draw)
        bd.set_out420p(to420)
        outs, fr = bd.decode([stream, stream, stream], fb, n)
        assert fr == [n] * 3
        assert all(hashlib.md5(o.tobytes()).hexdigest() == want for o in outs)
        d_out = torch.zeros(fb * n, dtype=torch.uint8, device="cuda")
        sb = np.frombuffer(stream, dtype=np.uint8)
        rc, fr = bd.decode_ptrs([sb.ctypes.data], None, [len(sb)], [d_out.data_ptr()], [fb * n], 1)
        torch.cuda.synchronize()
        bd.close()
        assert rc == 0 and fr == [n]
        assert hashlib.md5(d_out.cpu().numpy().tobytes()).hexdigest() == want
