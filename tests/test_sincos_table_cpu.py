"""quasi_b200/csrc/sincos_table.hpp (the kernels' sin/cos, device_common.cuh:sincos_round_once)
on the host: for every phi = 2*pi*r1 the shader's RNG can produce it must
return the bits the oracle defines, (float)sin((double)phi) and (float)cos((double)phi)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_table_file_is_what_the_generator_writes():
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "make_sincos_table.py"), "--check"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_every_drawable_angle_gets_the_oracles_bits(tmp_path):
    exe = tmp_path / "sincos_table_check"
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread",
                    f"-I{ROOT / 'quasi_b200' / 'csrc'}", str(ROOT / "tests" / "sincos_table_check.cpp"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe), "1"], capture_output=True, text=True)   # all 83 886 081 of them, a few seconds
    assert out.returncode == 0, out.stdout[-2000:]
    words = out.stdout.split()
    assert int(words[words.index("tested") + 1]) == 5 * 2**24 + 1
    assert int(words[words.index("sin_mismatches") + 1]) + int(words[words.index("cos_mismatches") + 1]) == 0
