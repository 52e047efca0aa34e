"""CPU: the whole-buffer DEFLATE decoder of the ingestion path (skani_b200/cli/fast_inflate.hpp) against zlib -- every
compression level and strategy, stored / fixed / dynamic blocks, multi-member gzip, trailing bytes, truncated and corrupted
streams (rejected or decoded like zlib, never garbage), the gzip fixtures of tests/golden/, and the block-parallel two-pass
decoding of one member (equal to the text, or declined: binary data, several members, corruption).  See tests/emu/emu_inflate.cpp."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_inflate_matches_zlib(tmp_path):
    exe = str(tmp_path / "emu_inflate")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_inflate.cpp"), "-lz"])
    fixtures = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.gz")))
    assert fixtures
    out = subprocess.run([exe] + fixtures, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " cases, 0 failures" in out.stdout and "rates" in out.stdout and "parallel gunzip" in out.stdout
