"""The three synthetic generators (numpy spec, C twin) agree; md5 from SURVEY.md Appendix C."""
import hashlib

import numpy as np

import dsvlibs as L
import synth


def test_c_matches_numpy_small():
    for (w, h, fmt, seed, cut) in [(64, 48, "420", 5, 0), (96, 64, "444", 9, 2), (80, 64, "422", 3, 1),
                                   (128, 32, "411", 7, 0)]:
        a = synth.sequence(w, h, 4, fmt, seed, cut)
        b = L.synth_sequence(w, h, fmt, 4, seed, cut)
        assert a == b.tobytes(), (w, h, fmt)


def test_cif_md5_prefix():
    # first 6 frames + frames 149..151 of the CIF vector (full-vector md5 is checked in test_golden_streams)
    a = synth.sequence(352, 288, 2, "420", 1, 150)
    b = L.synth_sequence(352, 288, "420", 2, 1, 150)
    assert a == b.tobytes()
    a = synth.sequence(352, 288, 2, "420", 1, 150, start=149)
    b = L.synth_sequence(352, 288, "420", 2, 1, 150, start=149)
    assert a == b.tobytes()


def test_cif_full_md5():
    y = L.synth_sequence(352, 288, "420", 300, 1, 150)
    assert hashlib.md5(y.tobytes()).hexdigest() == "ee3a62a3077c78292ca3b739330f68ef"
