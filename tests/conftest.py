"""pytest configuration: the `gpu` marker and the CPU-checker fixtures."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import oracle_lib  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
LAYERS = ("beauty", "albedo", "normal", "depth")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    oracle_lib.build_oracle(ref=True)
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference(oracle):
    if not oracle_lib.Reference.available():
        pytest.skip("oracle/_ref/libquasi_ref.so not built (needs /root/reference at build time)")
    return oracle_lib.Reference()


@pytest.fixture(scope="session")
def golden():
    class G:
        dir = GOLDEN
        scene = json.loads((GOLDEN / "cornell_box_scene.json").read_text())
        kat = json.loads((GOLDEN / "pcg_kat.json").read_text())
        checksums = json.loads((GOLDEN / "ref_checksums.json").read_text())

        @staticmethod
        def npz(name):
            return dict(np.load(GOLDEN / name))

        @staticmethod
        def uniforms(tag):
            raw = bytes.fromhex(G.scene[tag]["uniforms_hex"])
            return np.frombuffer(raw, dtype=oracle_lib.uniforms_dtype)[0].copy()

    return G
