"""Single-sequence sharding (dsvb_encode_long / dsvb_decode_long / dsvb_multi_*, csrc/host/long.cpp): ONE sequence
spread over the lanes of a GPU by I-delimited chains gives exactly the reference's bytes -- the CIF-300 golden with
its forced I picture at 150 (after which a fresh-encoder-per-GOP scheme diverges), HD goldens,
host / device / streaming inputs, ABR falling back to one lane."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")))


def _md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.mark.parametrize("lanes", [1, 4, 25])
def test_cif300_golden_sharded_by_chains(gpu, lanes):
    g = GOLD["cif_gop12"]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], g["cut"])
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"])
    be = L.BatchEncoder(gpu, cfg, lanes)
    stream, info = be.encode_long(yuv, n)
    be.close()
    assert len(stream) == g["dsv_len"] and _md5(stream) == g["dsv_md5"], info
    assert info[0] == 26 and info[1] == 1  # 25 GOP starts + the forced I picture at the cut
    bd = L.BatchDecoder(gpu, lanes)
    out, fr = bd.decode_long(stream, fb, n)
    bd.close()
    assert fr == n and _md5(out.tobytes()) == g["dec_md5"]


@pytest.mark.parametrize("name,lanes", [("hd_gop12", 2), ("hd_gop12_qp50", 3), ("hd_gop0", 8), ("qcif_gop12_444", 5),
                                        ("qcif_gop12_411", 4), ("w854_gop6", 2)])
def test_goldens_sharded_by_chains(gpu, name, lanes):
    g = GOLD[name]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], g["cut"])
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"])
    be = L.BatchEncoder(gpu, cfg, lanes)
    stream, info = be.encode_long(yuv, n)
    be.close()
    assert _md5(stream) == g["dsv_md5"], info
    bd = L.BatchDecoder(gpu, lanes)
    out, fr = bd.decode_long(stream, fb, n)
    bd.close()
    assert fr == n and _md5(out.tobytes()) == g["dec_md5"]


def test_long_input_paths_agree(gpu):
    """device-resident input, host input through the device cache, host input streamed twice (cache disabled),
    pinned host input: same bytes; a stream buffer that is too small is reported"""
    import torch
    w, h, fmt, n = 352, 288, "420", 40
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, n, 11, 17)
    cfg = L.make_cfg(w, h, fmt, gop=12)
    want = gpu.encode_sequence(cfg, yuv, n)[0]
    be = L.BatchEncoder(gpu, cfg, 6)
    assert be.encode_long(yuv, n)[0] == want
    os.environ["DSV_LONG_CACHE_MB"] = "0"
    try:
        assert be.encode_long(yuv, n)[0] == want
        pinned = torch.from_numpy(yuv).pin_memory()
        out = np.zeros(len(yuv) * 2, dtype=np.uint8)
        rc, ln, _ = be.encode_long_ptr(pinned.data_ptr(), n, 0, out.ctypes.data, len(out))
        assert rc == 0 and out[:ln].tobytes() == want
    finally:
        del os.environ["DSV_LONG_CACHE_MB"]
    d = torch.from_numpy(yuv).cuda()
    out = np.zeros(len(yuv) * 2, dtype=np.uint8)
    rc, ln, _ = be.encode_long_ptr(d.data_ptr(), n, 1, out.ctypes.data, len(out))
    assert rc == 0 and out[:ln].tobytes() == want
    guard = np.full(8192, 0xAB, dtype=np.uint8)
    rc, ln, _ = be.encode_long_ptr(d.data_ptr(), n, 1, guard.ctypes.data, 4096)
    assert rc == -1 and ln == -1 and (guard[4096:] == 0xAB).all()
    be.close()


def test_long_abr_runs_serial(gpu, ref):
    """ABR depends on every previous packet size: the long entry runs it on one lane and still matches"""
    w, h, fmt, n = 352, 288, "420", 20
    yuv = L.synth_sequence(w, h, fmt, n, 3, 0)
    cfg = L.make_cfg(w, h, fmt, gop=12, rc_mode=1, bitrate=400000, quality=L.qp_to_quality(60))
    want = ref.encode_sequence(cfg, yuv, n)[0]
    be = L.BatchEncoder(gpu, cfg, 4)
    got, info = be.encode_long(yuv, n)
    be.close()
    assert got == want and info[3] == 1


def test_long_no_scd_intra_share_fallback(gpu, ref):
    """scene-change detection off: the cut is caught by the motion search's intra share (dsv_encoder.c:246-253), the
    picture becomes a forced I picture, in the serial API and in the sharded one"""
    w, h, fmt, n = 352, 288, "420", 16
    yuv = L.synth_sequence(w, h, fmt, n, 5, 9)
    cfg = L.make_cfg(w, h, fmt, gop=12, do_scd=0)
    want, pk, _ = ref.encode_sequence(cfg, yuv, n)
    assert gpu.encode_sequence(cfg, yuv, n)[0] == want
    be = L.BatchEncoder(gpu, cfg, 3)
    got, info = be.encode_long(yuv, n)
    be.close()
    assert got == want
    assert info[1] >= 1, "the cut was expected to exceed the intra share threshold"


def test_multi_gpu_object(gpu):
    """dsvb_multi_*: all visible GPUs (and, with one GPU, two engines on it) behind one object"""
    import torch
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    g = GOLD["cif_gop12"]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], 120
    sub = L.SUBSAMP[fmt]
    fb = L.frame_bytes(w, h, sub)
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], 50)
    cfg = L.make_cfg(w, h, fmt, gop=g["gop"], qp=g["qp"])
    want = gpu.encode_sequence(cfg, yuv, n)[0]
    mg = L.MultiGpu(gpu, cfg, 4, devices)
    got, info = mg.encode_long(yuv, n)
    assert got == want and info[3] == len(devices)
    out, fr = mg.decode_long(want, fb, n)
    assert fr == n and np.array_equal(out, gpu.decode_stream(want, w, h, sub, n)[1])
    seqs = [L.synth_sequence(w, h, fmt, 5, 70 + i, 0) for i in range(5)]
    streams = mg.encode(seqs, 5)
    assert streams == [gpu.encode_sequence(cfg, s, 5)[0] for s in seqs]
    outs, frs = mg.decode(streams, fb, 5)
    mg.close()
    assert frs == [5] * 5
    for o, s in zip(outs, streams):
        assert np.array_equal(o, gpu.decode_stream(s, w, h, sub, 5)[1])
