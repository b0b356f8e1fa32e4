"""include/dsv1_b200.h lays out every public struct exactly like the reference headers.

tests/golden/abi_layout.txt was produced by tools/abi_probe.c compiled against the
reference headers (gcc -I/root/reference tools/abi_probe.c); here the same probe is
compiled against include/compat and must print the same sizes/offsets."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe(inc, tmp_path, name):
    exe = str(tmp_path / name)
    subprocess.run(["gcc", "-I" + inc, os.path.join(ROOT, "tools", "abi_probe.c"), "-o", exe], check=True)
    return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


def test_layout_matches_golden(tmp_path):
    mine = _probe(os.path.join(ROOT, "include", "compat"), tmp_path, "mine")
    gold = open(os.path.join(ROOT, "tests", "golden", "abi_layout.txt")).read()
    assert mine == gold


def test_layout_matches_reference_headers(tmp_path):
    if not os.path.exists("/root/reference/dsv.h"):
        import pytest
        pytest.skip("reference tree not present")
    assert _probe("/root/reference", tmp_path, "ref") == _probe(os.path.join(ROOT, "include", "compat"), tmp_path, "mine")
