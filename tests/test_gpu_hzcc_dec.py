"""CUDA HZCC decoder (hzcc_dec.cu: parallel bit-FSM parse + scatter) through the C ABI vs the checkers."""
import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

PLANES = [(16, 16), (120, 68), (960, 540), (428, 240), (352, 288), (136, 68), (854, 480), (1920, 1080)]


def sparse_plane(rng, cw, ch, dens, scale=400):
    co = (rng.laplace(0, scale, size=(ch, cw)) * (rng.random((ch, cw)) < dens)).astype(np.int32)
    co[0, 0] = int(rng.integers(-30000, 30000))
    return co


@pytest.mark.parametrize("dims", PLANES)
def test_decode_plane(gpu, port, dims):
    cw, ch = dims
    rng = np.random.default_rng(cw * 5 + ch)
    for isP in (0, 1):
        for c in (0, 1):
            for q in (5, 313, 2047):
                nbh, nbv = int(rng.integers(1, 31)), int(rng.integers(1, 24))
                stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
                co = sparse_plane(rng, cw, ch, float(rng.choice([0.3, 0.02, 0.001])))
                s, _ = port.encode_plane(co, q, isP, c, stable, nbh, nbv)
                want = port.decode_plane(s, cw, ch, q, isP, c, stable, nbh, nbv)
                got = gpu.decode_plane(s, cw, ch, q, isP, c, stable, nbh, nbv)
                assert np.array_equal(want, got)


def test_empty_plane(gpu, port):
    stable = np.zeros(6, dtype=np.uint8)
    z = np.zeros((48, 64), dtype=np.int32)
    s, _ = port.encode_plane(z, 313, 0, 0, stable, 3, 2)
    assert np.array_equal(gpu.decode_plane(s, 64, 48, 313, 0, 0, stable, 3, 2), z)


def test_truncated_stream(gpu, port):
    """plen shorter than the coded data: decoding stops where the reference's byte-pointer check does."""
    rng = np.random.default_rng(3)
    stable = rng.integers(0, 4, size=12, dtype=np.uint8)
    co = sparse_plane(rng, 176, 144, 0.2)
    s, _ = port.encode_plane(co, 313, 1, 0, stable, 4, 3)
    for cut in (len(s) // 2, len(s) // 3, 12):
        t = s.copy()
        t[:4] = np.frombuffer(int(cut).to_bytes(4, "big"), dtype=np.uint8)
        assert np.array_equal(port.decode_plane(t, 176, 144, 313, 1, 0, stable, 4, 3),
                              gpu.decode_plane(t, 176, 144, 313, 1, 0, stable, 4, 3))


def test_uhd_vs_reference(gpu, ref):
    rng = np.random.default_rng(6)
    stable = rng.integers(0, 4, size=60 * 34, dtype=np.uint8)
    co = sparse_plane(rng, 3840, 2160, 0.25)
    s, _ = ref.encode_plane(co, 313, 0, 0, stable, 60, 34)
    assert np.array_equal(ref.decode_plane(s, 3840, 2160, 313, 0, 0, stable, 60, 34),
                          gpu.decode_plane(s, 3840, 2160, 313, 0, 0, stable, 60, 34))
