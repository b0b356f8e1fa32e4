"""Size / format / quantiser matrix through the drop-in API vs the unmodified reference:
every block size class (16/24/32/48/64 per dimension), all four subsamplings, odd chroma sizes, the smallest legal
picture, dimensions that are not multiples of 8, qp 0 / 50 / 100."""
import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, fmt, frames, gop, qp
    (16, 16, "420", 4, 12, 85),        # smallest picture the library accepts: one block
    (18, 22, "444", 4, 12, 85),        # not a multiple of anything useful
    (64, 48, "411", 5, 12, 50),
    (358, 202, "420", 5, 12, 85),      # blocks 24x16, odd chroma (179x101)
    (360, 200, "422", 5, 4, 100),
    (708, 358, "420", 4, 12, 0),       # blocks 32x24, qp 0 (coarsest quantiser)
    (1030, 360, "420", 3, 12, 85),     # blocks 48x24
    (1284, 724, "420", 3, 12, 50),     # blocks 64x32, width not a multiple of 8
    (1300, 1026, "444", 2, 12, 85),    # blocks 64x48
    (854, 480, "411", 3, 12, 85),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_%s_gop%d_qp%d" % (c[0], c[1], c[2], c[4], c[5]))
def test_matrix_vs_reference(gpu, ref, case):
    w, h, fmt, n, gop, qp = case
    sub = L.SUBSAMP[fmt]
    if min(w, h) >= 64:
        yuv = L.synth_sequence(w, h, fmt, n, w + h, 0)
    else:   # the synthetic scene needs room for its moving object: plain noise + drift for tiny pictures
        rng = np.random.default_rng(w * h)
        fb = L.frame_bytes(w, h, sub)
        base = rng.integers(0, 256, size=fb, dtype=np.uint8)
        yuv = np.concatenate([np.roll(base, 3 * t) for t in range(n)])
    cfg = L.make_cfg(w, h, fmt, gop=gop, qp=qp)
    sa, pa, _ = ref.encode_sequence(cfg, yuv, n)
    sb, pb, _ = gpu.encode_sequence(cfg, yuv, n)
    assert pa == pb
    assert sa == sb
    na, da, _, _ = ref.decode_stream(sa, w, h, sub, n)
    nb, db, _, _ = gpu.decode_stream(sa, w, h, sub, n)
    assert na == nb == n
    assert np.array_equal(da, db)
