"""CUDA quantiser + HZCC encoder (quant.cuh, hzcc_enc.cu) through the C ABI vs the checkers:
stream bytes and dequantised write-back, byte/integer exact."""
import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

PLANES = [(16, 16), (120, 68), (960, 540), (428, 240), (352, 288), (176, 144), (136, 68), (854, 480), (240, 136),
          (1920, 1080)]


def sparse_plane(rng, cw, ch, dens, scale=400):
    co = (rng.laplace(0, scale, size=(ch, cw)) * (rng.random((ch, cw)) < dens)).astype(np.int32)
    co[0, 0] = int(rng.integers(-30000, 30000))
    return co


@pytest.mark.parametrize("dims", PLANES)
def test_encode_plane(gpu, port, dims):
    cw, ch = dims
    rng = np.random.default_rng(cw * 7 + ch)
    for isP in (0, 1):
        for c in (0, 1):
            for q in (5, 313, 900, 2047):
                nbh, nbv = int(rng.integers(1, 31)), int(rng.integers(1, 24))
                stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
                co = sparse_plane(rng, cw, ch, float(rng.choice([0.3, 0.02, 0.001])))
                sa, ca = port.encode_plane(co, q, isP, c, stable, nbh, nbv)
                sb, cb = gpu.encode_plane(co, q, isP, c, stable, nbh, nbv)
                assert np.array_equal(ca, cb)
                assert np.array_equal(sa, sb)


def test_empty_plane(gpu, port):
    stable = np.zeros(6, dtype=np.uint8)
    z = np.zeros((48, 64), dtype=np.int32)
    sa, _ = port.encode_plane(z, 313, 0, 0, stable, 3, 2)
    sb, _ = gpu.encode_plane(z, 313, 0, 0, stable, 3, 2)
    assert np.array_equal(sa, sb)


def test_uhd_vs_reference(gpu, ref):
    rng = np.random.default_rng(5)
    stable = rng.integers(0, 4, size=60 * 34, dtype=np.uint8)
    co = sparse_plane(rng, 3840, 2160, 0.25)
    sa, ca = ref.encode_plane(co, 313, 0, 0, stable, 60, 34)
    sb, cb = gpu.encode_plane(co, 313, 0, 0, stable, 60, 34)
    assert np.array_equal(ca, cb) and np.array_equal(sa, sb)


@pytest.mark.parametrize("dims", [(352, 288, 352, 288), (959, 539, 960, 540), (427, 240, 428, 240), (1920, 1080, 1920, 1080),
                                  (214, 120, 214, 120), (107, 60, 108, 60), (240, 135, 240, 136)])
def test_fused_quant_equals_separate(gpu, port, dims):
    """SBT epilogue quantiser == transform followed by encode_plane's in-place write-back."""
    pw, ph, cw, ch = dims
    rng = np.random.default_rng(pw)
    y, x = np.mgrid[0:ph, 0:cw + 2]
    pix = np.clip(128 + 70 * np.sin(x / 5.0) * np.cos(y / 6.0) + rng.integers(-20, 21, size=(ph, cw + 2)), 0, 255).astype(np.uint8)
    bw, bh, nbh, nbv = L.block_dims(pw if pw == cw else pw * 2, ph if ph == ch else ph * 2)
    for isP in (0, 1):
        for c in (0, 1):
            stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
            raw = port.fwd_sbt(pix, pw, ph, cw, ch, isP)
            _, want = port.encode_plane(raw, 313, isP, c, stable, nbh, nbv)
            got, _ = L.fwd_sbt_q(gpu, pix, pw, ph, cw, ch, isP, c, 313, stable, nbh, nbv)
            assert np.array_equal(want, got)
