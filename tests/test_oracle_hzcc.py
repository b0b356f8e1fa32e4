"""oracle/dsv1_port.c quantiser + HZCC coder == unmodified reference (hzcc.c, bs.c): stream bytes,
in-place dequantised write-back and decoded planes, incl. the double-visited rows of 960x540."""
import numpy as np
import pytest

import dsvlibs as L

PLANES = [(120, 68), (960, 540), (428, 240), (352, 288), (176, 144), (136, 68), (16, 16), (854, 480), (240, 136)]


def sparse_plane(rng, cw, ch, dens, scale=400):
    co = (rng.laplace(0, scale, size=(ch, cw)) * (rng.random((ch, cw)) < dens)).astype(np.int32)
    co[0, 0] = int(rng.integers(-30000, 30000))
    return co


@pytest.mark.parametrize("dims", PLANES)
def test_encode_decode(ref, port, dims):
    cw, ch = dims
    rng = np.random.default_rng(cw * 7 + ch)
    for isP in (0, 1):
        for c in (0, 1):
            for q in (5, 313, 500, 900, 2047):
                nbh, nbv = int(rng.integers(1, 31)), int(rng.integers(1, 24))
                stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
                if q == 500:
                    stable[:] = 0
                co = sparse_plane(rng, cw, ch, float(rng.choice([0.3, 0.02, 0.001])))
                sa, ca = ref.encode_plane(co, q, isP, c, stable, nbh, nbv)
                sb, cb = port.encode_plane(co, q, isP, c, stable, nbh, nbv)
                assert np.array_equal(sa, sb)
                assert np.array_equal(ca, cb)
                assert np.array_equal(ref.decode_plane(sa, cw, ch, q, isP, c, stable, nbh, nbv),
                                      port.decode_plane(sa, cw, ch, q, isP, c, stable, nbh, nbv))


def test_empty_and_single(ref, port):
    stable = np.zeros(6, dtype=np.uint8)
    for cw, ch in [(64, 48), (960, 540)]:
        z = np.zeros((ch, cw), dtype=np.int32)
        sa, _ = ref.encode_plane(z, 313, 0, 0, stable, 3, 2)
        sb, _ = port.encode_plane(z, 313, 0, 0, stable, 3, 2)
        assert np.array_equal(sa, sb)
        assert np.array_equal(port.decode_plane(sb, cw, ch, 313, 0, 0, stable, 3, 2), z)


def test_double_visit_probe(port):
    """SURVEY.md Appendix B-1 probe values (obtained from the reference): 960x540 plane, one coefficient
    10000 at (5,135), q=313, I frame: nruns=2, write-back 9880; at y=134: nruns=1, 9859."""
    stable = np.zeros(690, dtype=np.uint8)
    for y, nruns, wb in [(135, 2, 9880), (134, 1, 9859)]:
        z = np.zeros((540, 960), dtype=np.int32)
        z[y, 5] = 10000
        s, c = port.encode_plane(z, 313, 0, 1, stable, 30, 23)
        # layout: plen u32 | SEG(0) = '1' + pad -> 1 byte | nruns u32
        assert int.from_bytes(bytes(s[5:9]), "big") == nruns
        assert c[y, 5] == wb


def test_truncated_stream(ref, port):
    rng = np.random.default_rng(3)
    stable = rng.integers(0, 4, size=12, dtype=np.uint8)
    co = sparse_plane(rng, 176, 144, 0.2)
    s, _ = ref.encode_plane(co, 313, 1, 0, stable, 4, 3)
    for cut in (len(s) // 2, len(s) // 3, 12):
        t = s.copy()
        t[:4] = np.frombuffer(int(cut).to_bytes(4, "big"), dtype=np.uint8)   # lie about plen
        assert np.array_equal(ref.decode_plane(t, 176, 144, 313, 1, 0, stable, 4, 3),
                              port.decode_plane(t, 176, 144, 313, 1, 0, stable, 4, 3))
