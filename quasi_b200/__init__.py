"""quasi_b200: a B200-native CUDA rendering backend for the Quasi renderer.

The product is `quasi_b200/lib/libquasi_cuda.so`, a drop-in backend plugin
behind Quasi's C ABI (include/quasi_plugin_abi.h). This package holds the CUDA
sources (`csrc/`), the C++ headless host (`host/`) and a thin ctypes mirror of
the reference's plugin loader and GPU context used by the tests and bench.py.
Nothing here computes pixels on the CPU: if the library is missing, loading
fails loudly.
"""
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
REPO_ROOT = PACKAGE_DIR.parent
import os as _os

# QUASI_PLUGIN_OVERRIDE: measure another build of the backend (tools/build_variant.sh) with the same host code
PLUGIN_PATH = Path(_os.environ["QUASI_PLUGIN_OVERRIDE"]) if _os.environ.get("QUASI_PLUGIN_OVERRIDE") else \
    PACKAGE_DIR / "lib" / "libquasi_cuda.so"

from .abi import (  # noqa: E402,F401
    Q_AOV_BEAUTY, Q_AOV_ALBEDO, Q_AOV_NORMAL, Q_AOV_DEPTH, Q_AOV_COUNT,
    Q_GPU_BACKEND_CUDA, Q_GPU_BACKEND_METAL, K_PLUGIN_ABI_VERSION,
)
from .loader import dynamic_library, library_error, loader, loader_error  # noqa: E402,F401
from .cuda_context import cuda_context, orbit_camera  # noqa: E402,F401
