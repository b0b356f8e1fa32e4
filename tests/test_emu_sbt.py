"""Kernel-logic check in the GPU-less container: the SAME .cu sources compiled for the test-only CPU
emulator (tests/emu) vs the checkers.  Small sizes only; the real parity tests are the -m gpu ones."""
import os
import subprocess

import numpy as np
import pytest

import dsvlibs as L


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", L.PKG, "emu"], check=True, stdout=subprocess.DEVNULL)
    return L.emu()


@pytest.mark.parametrize("dims", [(16, 16, 16, 16), (120, 68, 120, 68), (135, 67, 136, 68), (427, 240, 428, 240), (384, 192, 384, 192)])
@pytest.mark.parametrize("isP", [0, 1])
def test_sbt(emu, port, dims, isP):
    pw, ph, cw, ch = dims
    rng = np.random.default_rng(pw + isP)
    pix = rng.integers(0, 256, size=(ph, cw + 2), dtype=np.uint8)
    a = port.fwd_sbt(pix, pw, ph, cw, ch, isP)
    assert np.array_equal(a, emu.fwd_sbt(pix, pw, ph, cw, ch, isP))
    co = (a // 7) * 7
    for c in (0, 1):
        assert np.array_equal(port.inv_sbt(co, 313, isP, c, pw, ph), emu.inv_sbt(co, 313, isP, c, pw, ph))


@pytest.mark.parametrize("dims", [(120, 68, 120, 68), (427, 240, 428, 240), (384, 192, 384, 192)])
def test_inverse_sparse(emu, port, dims):
    """Mostly-zero coefficient planes (what P pictures look like): flat LL areas next to isolated values."""
    from test_gpu_sbt import sparse_coefs
    pw, ph, cw, ch = dims
    rng = np.random.default_rng(pw)
    for density in (0.0, 0.002):
        co = sparse_coefs(rng, cw, ch, density)
        for c in (0, 1):
            assert np.array_equal(port.inv_sbt(co, 313, 1, c, pw, ph), emu.inv_sbt(co, 313, 1, c, pw, ph)), (density, c)
