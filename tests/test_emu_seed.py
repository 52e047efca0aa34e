"""Host emulation of the seeding kernels' per-unit arithmetic (skani_b200/csrc/sk_core.cuh, both the reference-shaped and
the tuned pass-mask functions) against the oracle's AVX2-semantics seeder: 400 random contigs with N runs at quarter-lane
boundaries, lowercase, IUPAC codes, every (len-20) mod 4, c in {1,10,30,125,200}, k in {13,15,16}."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unit_arithmetic_matches_oracle(tmp_path):
    exe = str(tmp_path / "emu_seed")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_seed.cpp"),
                           os.path.join(ROOT, "oracle", "skani_oracle.cpp"), "-lz"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "400 cases, 0 failures" in out.stdout
