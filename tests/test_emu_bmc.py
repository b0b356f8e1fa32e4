"""Streaming motion compensation (bmc.cu) on the test-only CPU emulator vs the unmodified reference
(dsv_sub_pred / dsv_add_pred, bmc.c:318-346): every subsampling, block widths that are not multiples of 16
(strips cut by block edges, 4:1:1 chroma blocks 6 wide), odd picture sizes, all four half-pel phases, vectors
that clamp into the border, whole and partial intra blocks.  The same cases run on the device in test_gpu_motion.py."""
import subprocess

import numpy as np
import pytest

import dsvlibs as L

CASES = [
    (112, 96, "444", None),
    (100, 70, "420", None),
    (103, 75, "411", (24, 16)),  # chroma blocks 6 wide; a 1-wide edge block would divide by zero in the reference (bmc.c:189)
    (90, 66, "422", (20, 28)),
    (132, 52, "420", (64, 48)),
    (75, 49, "420", (36, 20)),
]


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", L.PKG, "emu"], check=True, stdout=subprocess.DEVNULL)
    return L.emu()


def random_field(rng, nblk, amp=110, intra=0.3):
    mv = np.zeros(nblk, dtype=L.MV_DTYPE)
    mv["x"] = rng.integers(-amp, amp + 1, size=nblk).astype(np.int16)
    mv["y"] = rng.integers(-amp, amp + 1, size=nblk).astype(np.int16)
    mv["mode"] = (rng.random(nblk) < intra).astype(np.uint8)
    mv["submask"] = np.where(mv["mode"] == 1, rng.integers(1, 16, size=nblk), 0).astype(np.uint8)
    return mv


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d-%s-%s" % (c[0], c[1], c[2], "auto" if c[3] is None else "%dx%d" % c[3]))
def test_bmc_strips(emu, ref, case):
    w, h, fmt, blk = case
    sub = L.SUBSAMP[fmt]
    if blk is None:
        blk = L.block_dims(w, h)
    else:
        blk = (blk[0], blk[1], (w + blk[0] - 1) // blk[0], (h + blk[1] - 1) // blk[1])
    rng = np.random.default_rng(w * 131 + h)
    n = L.frame_bytes(w, h, sub)
    fr = rng.integers(0, 256, size=n, dtype=np.uint8)
    fs = rng.integers(0, 256, size=n, dtype=np.uint8)
    for amp, intra in ((110, 0.3), (5, 0.0)):
        mv = random_field(rng, blk[2] * blk[3], amp, intra)
        pa, ra = ref.sub_pred(mv, w, h, sub, fs, fr, blk)
        pb, rb = emu.sub_pred(mv, w, h, sub, fs, fr, blk)
        assert np.array_equal(pa, pb), "prediction differs"
        assert np.array_equal(ra, rb), "residual differs"
        assert np.array_equal(ref.add_pred(mv, w, h, sub, ra, fr, blk), emu.add_pred(mv, w, h, sub, ra, fr, blk))
