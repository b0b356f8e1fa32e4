"""HZCC encoder passes (warp-per-chunk scan / prefix / pack, hzcc_enc.cu) on the test-only CPU emulator vs the plain-C
port: stream bytes and dequantised write-back.  Small planes incl. odd sub-band sizes (double-visited positions),
widths that are not multiples of 4 (the 16-byte zero test falls back to single positions), dense and very sparse
planes, planes of more than one chunk.  The same cases at full size run on the device in test_gpu_hzcc_enc.py."""
import subprocess

import numpy as np
import pytest

import dsvlibs as L


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", L.PKG, "emu"], check=True, stdout=subprocess.DEVNULL)
    return L.emu()


@pytest.mark.parametrize("dims", [(16, 16), (120, 68), (136, 68), (214, 120), (108, 60), (240, 136), (90, 50)])
def test_encode_plane_emu(emu, port, dims):
    cw, ch = dims
    rng = np.random.default_rng(cw * 7 + ch)
    for isP in (0, 1):
        for c in (0, 1):
            for q, dens in ((5, 0.3), (313, 0.02), (900, 0.001)):
                nbh, nbv = int(rng.integers(1, 8)), int(rng.integers(1, 6))
                stable = rng.integers(0, 4, size=nbh * nbv, dtype=np.uint8)
                co = (rng.laplace(0, 400, size=(ch, cw)) * (rng.random((ch, cw)) < dens)).astype(np.int32)
                co[0, 0] = int(rng.integers(-30000, 30000))
                sa, ca = port.encode_plane(co, q, isP, c, stable, nbh, nbv)
                sb, cb = emu.encode_plane(co, q, isP, c, stable, nbh, nbv)
                assert np.array_equal(ca, cb)
                assert np.array_equal(sa, sb)


def test_empty_plane_emu(emu, port):
    stable = np.zeros(6, dtype=np.uint8)
    z = np.zeros((48, 64), dtype=np.int32)
    sa, _ = port.encode_plane(z, 313, 0, 0, stable, 3, 2)
    sb, _ = emu.encode_plane(z, 313, 0, 0, stable, 3, 2)
    assert np.array_equal(sa, sb)
