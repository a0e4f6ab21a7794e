#!/usr/bin/env python
"""Smallest possible GPU check of QUASI_CUDA_BATCH: 10 frames at 211x119 with batches of 4
(two full launches and a partial one flushed by the readback) against the CPU oracle, then
the same frames launch by launch. Prints one line per form."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

t0 = time.time()
import numpy as np  # noqa: E402

from oracle_lib import Oracle, bits  # noqa: E402
from quasi_b200.headless import headless_host  # noqa: E402

w, h, frames = 211, 119, 10
oracle = Oracle()
want, rays = oracle.render_frames(oracle.scene(w, h), w, h, frames)
for batch in (4, 1):
    host = headless_host(w, h, env={"QUASI_CUDA_BATCH": batch})
    host.render(frames)
    got, stats = host.readback_aov(), host.stats()
    host.close()
    bad = sum(int((bits(g) != bits(wv)).any(axis=-1).sum()) for g, wv in zip(got, want))
    print(f"batch {batch}: {bad} differing pixels, rays {stats.rays} (oracle {rays}), "
          f"{stats.kernel_launches} launches, t+{time.time() - t0:.1f}s", flush=True)
