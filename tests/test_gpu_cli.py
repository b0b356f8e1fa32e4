"""The UNMODIFIED reference command-line program (dsv_main.c, compiled from /root/reference where it lies, linked
against libdsv1_b200.so instead of the reference's own objects; built by `make cli` in the build container) encodes and
decodes the CIF-300 golden through the plain dsv_enc / dsv_dec API: same .dsv bytes, same decoded YUV as the
reference CLI produced for tests/golden/streams.json -- including -out420p and the default ABR mode."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import dsvlibs as L

pytestmark = pytest.mark.gpu

CLI = os.path.join(L.ROOT, "tests", "_cli", "dsv1_b200_cli")
GOLD = json.load(open(os.path.join(L.ROOT, "tests", "golden", "streams.json")))
FMT_FLAG = {"444": 0, "422": 1, "420": 2, "411": 3}


def _md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


@pytest.mark.parametrize("name", ["cif_gop12", "qcif_gop12_444", "hd_gop12_qp50"])
def test_reference_cli_binary_on_this_library(tmp_path, name):
    if not os.path.exists(CLI):
        pytest.skip("tests/_cli/dsv1_b200_cli not built (needs the reference source tree at build time)")
    g = GOLD[name]
    w, h, fmt, n = g["w"], g["h"], g["fmt"], g["frames"]
    yuv = L.synth_sequence(w, h, fmt, n, g["seed"], g["cut"])
    assert hashlib.md5(yuv.tobytes()).hexdigest() == g["yuv_md5"]
    src, dsv, out = str(tmp_path / "in.yuv"), str(tmp_path / "out.dsv"), str(tmp_path / "dec.yuv")
    yuv.tofile(src)
    r = subprocess.run([CLI, "e", "-y", "-inp_" + src, "-out_" + dsv, "-w%d" % w, "-h%d" % h, "-fmt%d" % FMT_FLAG[fmt],
                        "-gop%d" % g["gop"], "-qp%d" % g["qp"], "-rc_mode1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.getsize(dsv) == g["dsv_len"] and _md5_file(dsv) == g["dsv_md5"]
    r = subprocess.run([CLI, "d", "-y", "-inp_" + dsv, "-out_" + out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert _md5_file(out) == g["dec_md5"]
