"""oracle/dsv1_port_motion.c (plain-C restatement of pyramid, HME and BMC) == unmodified reference
(frame.c, hme.c, bmc.c): every DSV_MV byte, every predicted / residual / reconstructed sample."""
import math

import numpy as np
import pytest

import dsvlibs as L

CASES = [
    (352, 288, "420", 1, 150, [(0, 1), (149, 150), (150, 151)]),
    (176, 144, "444", 4, 14, [(2, 3), (13, 14)]),
    (176, 144, "422", 5, 0, [(0, 1)]),
    (176, 144, "411", 6, 0, [(7, 8)]),
    (428, 240, "420", 7, 0, [(3, 4)]),
]


def pyr_levels(w, h):
    bw, bh, nbh, nbv = L.block_dims(w, h)
    lv = int(math.ceil(math.log2(min(w, h))))
    while (1 << lv) > max(nbh, nbv):
        lv -= 1
    return min(max(lv, 3), 5)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_%s" % (c[0], c[1], c[2]))
def test_motion_port_vs_reference(ref, port, case):
    w, h, fmt, seed, cut, pairs = case
    sub = L.SUBSAMP[fmt]
    lv = pyr_levels(w, h)
    rng = np.random.default_rng(w)
    for a, b in pairs:
        fr = L.synth_sequence(w, h, fmt, 1, seed, cut, start=a)
        fs = L.synth_sequence(w, h, fmt, 1, seed, cut, start=b)
        for la, lb in zip(ref.pyramid(fs, w, h, sub, lv), port.pyramid(fs, w, h, sub, lv)):
            assert np.array_equal(la, lb)
        pr, mr = ref.hme(fs, fr, w, h, sub, lv)
        pp, mp = port.hme(fs, fr, w, h, sub, lv)
        assert pr == pp
        assert [k for k in mr.dtype.names if not np.array_equal(mr[k], mp[k])] == []
        for trial in range(2):
            mv = mr.copy()
            if trial:
                mv["x"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
                mv["y"] = rng.integers(-110, 111, size=mv.shape).astype(np.int16)
                mv["mode"] = (rng.random(mv.shape) < 0.3).astype(np.uint8)
                mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=mv.shape), 0).astype(np.uint8)
            p1, r1 = ref.sub_pred(mv, w, h, sub, fs, fr)
            p2, r2 = port.sub_pred(mv, w, h, sub, fs, fr)
            assert np.array_equal(p1, p2) and np.array_equal(r1, r2)
            assert np.array_equal(ref.add_pred(mv, w, h, sub, r1, fr), port.add_pred(mv, w, h, sub, r1, fr))
