"""Host-side ASCII -> 2-bit + N-mask packer (skani_b200/csrc/host_pack.hpp, AVX2/BMI2 with a scalar path) produces exactly
the units pack_kernel produces on the device (sk::ascii_code semantics): every byte value, lengths 0..5000, unaligned
starts, no write past the end.  Prepared for the PCIe-saving input path; see tests/emu/emu_pack.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_packer_matches_device_semantics(tmp_path):
    exe = str(tmp_path / "emu_pack")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_pack.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "4000 cases, 0 failures" in out.stdout
