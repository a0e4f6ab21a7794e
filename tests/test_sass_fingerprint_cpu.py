"""The default kernels of quasi_b200/lib/libquasi_cuda.so must be, instruction for instruction, the
ones the GPU parity suite last ran on (profiles/verified_sass.json, tools/sass_fingerprint.py).
This round added several build-time options and test hooks after its last GPU minute; each was
required to leave the default build's SASS untouched, and nvcc turned out to be sensitive even to
statements that do nothing. After a deliberate kernel change: run `pytest -m gpu` on a B200, then
`python tools/sass_fingerprint.py --update`."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump from the CUDA toolkit")
def test_default_kernels_are_the_ones_verified_on_the_gpu():
    if not (ROOT / "quasi_b200" / "lib" / "libquasi_cuda.so").exists():
        pytest.skip("libquasi_cuda.so has not been built")
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_fingerprint.py"), "--check"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
