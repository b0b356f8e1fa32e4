"""CUDA pyramid / hierarchical motion estimation / block motion compensation (frame_ops.cu, hme.cu, bmc.cu)
through the kernel-level C ABI vs the unmodified reference (dsv_hme, dsv_sub_pred, dsv_add_pred): every
DSV_MV byte, every predicted / residual / reconstructed sample."""
import math

import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, fmt, seed, cut, frame pairs (ref, src)
    (352, 288, "420", 1, 150, [(0, 1), (10, 11), (149, 150), (150, 151), (200, 201)]),
    (176, 144, "444", 4, 14, [(2, 3), (13, 14)]),
    (176, 144, "422", 5, 0, [(0, 1)]),
    (176, 144, "411", 6, 0, [(7, 8)]),
    (428, 240, "420", 7, 0, [(3, 4)]),
    (854, 480, "420", 8, 0, [(1, 2)]),
    (1920, 1080, "420", 2, 0, [(3, 4), (0, 1)]),
    (1280, 720, "420", 12, 0, [(5, 6)]),
]


def pyr_levels(w, h):
    """dsv_encoder.c:602-613"""
    bw, bh, nbh, nbv = L.block_dims(w, h)
    lv = int(math.ceil(math.log2(min(w, h))))
    while (1 << lv) > max(nbh, nbv):
        lv -= 1
    return min(max(lv, 3), 5)


def mv_equal(a, b):
    return [k for k in a.dtype.names if not np.array_equal(a[k], b[k])]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_%s" % (c[0], c[1], c[2]))
def test_hme_and_bmc(gpu, ref, case):
    w, h, fmt, seed, cut, pairs = case
    sub = L.SUBSAMP[fmt]
    lv = pyr_levels(w, h)
    rng = np.random.default_rng(w + h)
    for a, b in pairs:
        fr = L.synth_sequence(w, h, fmt, 1, seed, cut, start=a)
        fs = L.synth_sequence(w, h, fmt, 1, seed, cut, start=b)
        for la, lb in zip(ref.pyramid(fs, w, h, sub, lv), gpu.pyramid(fs, w, h, sub, lv)):
            assert np.array_equal(la, lb)
        pr, mr = ref.hme(fs, fr, w, h, sub, lv)
        pg, mg = gpu.hme(fs, fr, w, h, sub, lv)
        assert mv_equal(mr, mg) == []
        assert pr == pg
        # BMC with the estimated field, then with exaggerated vectors and forced intra / partial masks.
        # |mv| <= 110 half-pels keeps every filter tap inside the reference's own allocation: beyond that the
        # reference reads heap bytes outside its frame, which nothing can reproduce.
        for trial in range(3):
            mv = mr.copy()
            if trial >= 1:
                mv["x"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
                mv["y"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
                mv["mode"] = (rng.random(mv.shape) < 0.3).astype(np.uint8)
                mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=mv.shape), 0).astype(np.uint8)
            pred_r, res_r = ref.sub_pred(mv, w, h, sub, fs, fr)
            pred_g, res_g = gpu.sub_pred(mv, w, h, sub, fs, fr)
            assert np.array_equal(pred_r, pred_g)
            assert np.array_equal(res_r, res_g)
            assert np.array_equal(ref.add_pred(mv, w, h, sub, res_r, fr), gpu.add_pred(mv, w, h, sub, res_r, fr))


def test_uhd444_hme(gpu, ref):
    w, h, fmt = 3840, 2160, "444"
    sub = L.SUBSAMP[fmt]
    fr = L.synth_sequence(w, h, fmt, 1, 3, 0, start=1)
    fs = L.synth_sequence(w, h, fmt, 1, 3, 0, start=2)
    lv = pyr_levels(w, h)
    pr, mr = ref.hme(fs, fr, w, h, sub, lv)
    pg, mg = gpu.hme(fs, fr, w, h, sub, lv)
    assert mv_equal(mr, mg) == [] and pr == pg
    pred_r, res_r = ref.sub_pred(mr, w, h, sub, fs, fr)
    pred_g, res_g = gpu.sub_pred(mr, w, h, sub, fs, fr)
    assert np.array_equal(pred_r, pred_g) and np.array_equal(res_r, res_g)


def test_dsv_hme_exported_interface(gpu, ref):
    """dsv_hme(DSV_HME *) itself (dsv_encoder.h:122-132): caller-built bordered pyramids in, dsv_alloc'd vector
    fields of every level out -- driven by tools/api_harness.c for both libraries."""
    for (w, h, fmt, lv) in [(176, 144, "420", 3), (90, 70, "444", 2), (640, 360, "422", 4), (1920, 1080, "420", 4)]:
        sub = L.SUBSAMP[fmt]
        fr = L.synth_sequence(w, h, fmt, 1, 4, 0, start=3)
        fs = L.synth_sequence(w, h, fmt, 1, 9 if w < 1000 else 4, 0, start=4)
        pr, mr = ref.hme_api(fs, fr, w, h, sub, lv)
        pg, mg = gpu.hme_api(fs, fr, w, h, sub, lv)
        assert pr == pg
        for k in mr.dtype.names:
            if k != "pad":
                assert np.array_equal(mr[k], mg[k]), (w, h, fmt, k)


STRIP_CASES = [
    (112, 96, "444", None), (100, 70, "420", None), (103, 75, "411", (24, 16)), (90, 66, "422", (20, 28)),
    (132, 52, "420", (64, 48)), (75, 49, "420", (36, 20)), (1920, 1080, "420", None), (1280, 720, "422", (44, 52)),
    (854, 480, "411", None), (3840, 2160, "444", None),
]


@pytest.mark.parametrize("case", STRIP_CASES, ids=lambda c: "%dx%d-%s-%s" % (c[0], c[1], c[2], "auto" if c[3] is None else "%dx%d" % c[3]))
def test_bmc_strips_random_fields(gpu, ref, case):
    """The streaming compensation on random frames and random vector fields (all half-pel phases, vectors clamping
    into the border, whole / partial intra blocks), block widths that are not multiples of 16, every subsampling."""
    w, h, fmt, blk = case
    sub = L.SUBSAMP[fmt]
    blk = L.block_dims(w, h) if blk is None else (blk[0], blk[1], (w + blk[0] - 1) // blk[0], (h + blk[1] - 1) // blk[1])
    rng = np.random.default_rng(w * 131 + h)
    n = L.frame_bytes(w, h, sub)
    fr = rng.integers(0, 256, size=n, dtype=np.uint8)
    fs = rng.integers(0, 256, size=n, dtype=np.uint8)
    for amp, intra in ((110, 0.3), (5, 0.0)):
        nblk = blk[2] * blk[3]
        mv = np.zeros(nblk, dtype=L.MV_DTYPE)
        mv["x"] = rng.integers(-amp, amp + 1, size=nblk).astype(np.int16)
        mv["y"] = rng.integers(-amp, amp + 1, size=nblk).astype(np.int16)
        mv["mode"] = (rng.random(nblk) < intra).astype(np.uint8)
        mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=nblk), 0).astype(np.uint8)
        pa, ra = ref.sub_pred(mv, w, h, sub, fs, fr, blk)
        pb, rb = gpu.sub_pred(mv, w, h, sub, fs, fr, blk)
        assert np.array_equal(pa, pb), "prediction differs"
        assert np.array_equal(ra, rb), "residual differs"
        assert np.array_equal(ref.add_pred(mv, w, h, sub, ra, fr, blk), gpu.add_pred(mv, w, h, sub, ra, fr, blk))
