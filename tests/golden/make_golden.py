#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container, where /root/reference exists:

    make -C oracle all ref && python tests/golden/make_golden.py

Everything written here comes from oracle/_ref/libquasi_ref.so, i.e. the
reference's own shader text (/root/reference/backends/metal/plugin.mm:26-351)
compiled with g++ against oracle/msl_shim/, and the reference's own
make_cornell_box() (/root/reference/src/quasi/scene/cornell_box.hpp:95-134).
The plain-C oracle and the CUDA backend are both tested against these files;
neither is used to produce them.

Outputs (all small):
  cornell_box_scene.json   the packed scene block for aspect 1 and 16/9, as hex
  pcg_kat.json             integer known-answer vectors
                           re-derived by running the shader's pcg_hash
  ref_48x48_f0.npz         one 1-spp sample, 4 layers, + tonemapped BGRA8
  ref_64x36_acc4.npz       4 accumulated frames (frame indices 0..3), 4 layers
  ref_64x36_shard.npz      frames 1,3,5 (first=1, stride=2): the sample-shard path
  ref_checksums.json       sha256 of the raw float32 bytes of full-size renders
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_lib import Reference, uniforms_dtype  # noqa: E402

LAYERS = ("beauty", "albedo", "normal", "depth")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    ref = Reference()

    # scene block
    scene = {}
    for name, (w, h) in {"aspect_1": (512, 512), "aspect_16_9": (1920, 1080)}.items():
        u = ref.scene(w, h)
        scene[name] = {
            "width": w,
            "height": h,
            "uniforms_hex": u.tobytes().hex(),
            "quad_count": int(u["quad_count"]),
            "light_index": int(u["light_index"]),
        }
    scene["scene_camera_aspect_1"] = [float(x) for x in ref.scene_camera(1.0)]
    (HERE / "cornell_box_scene.json").write_text(json.dumps(scene, indent=1) + "\n")

    # integer path
    def seed(px, py, f):
        return ref.pcg_hash((px + py * 1920 + f * 1920 * 1080) & 0xFFFFFFFF)

    kat = {
        "pcg_hash": {str(x): ref.pcg_hash(x) for x in (0, 1, 2, 0xFFFFFFFF, 1920, 2073600, 12345678)},
        "seed": [
            [px, py, f, seed(px, py, f)]
            for (px, py, f) in [
                (0, 0, 0), (1, 0, 0), (0, 1, 0), (959, 539, 0), (1919, 1079, 0), (0, 0, 1),
                (0, 0, 63), (3839, 2159, 1023), (0, 0, 2071), (0, 0, 2072),
            ]
        ],
        "draws": {},
    }
    for px, py, f in [(0, 0, 0), (1, 0, 0), (511, 511, 0), (1919, 1079, 63)]:
        states, floats = ref.random_floats(seed(px, py, f), 8)
        kat["draws"][f"{px},{py},{f}"] = {
            "states": states,
            "float_bits": [int(np.float32(x).view(np.uint32)) for x in floats],
        }
    (HERE / "pcg_kat.json").write_text(json.dumps(kat, indent=1) + "\n")

    # small images
    u = ref.scene(48, 48)
    layers = ref.render_sample(u, 48, 48)
    np.savez_compressed(
        HERE / "ref_48x48_f0.npz",
        **dict(zip(LAYERS, layers)),
        tonemap_bgra8=ref.tonemap_bgra8(layers[0]),
    )

    u = ref.scene(64, 36)
    np.savez_compressed(HERE / "ref_64x36_acc4.npz", **dict(zip(LAYERS, ref.render_frames(u, 64, 36, 4))))
    np.savez_compressed(
        HERE / "ref_64x36_shard.npz",
        **dict(zip(LAYERS, ref.render_frames(u, 64, 36, 3, frame_first=1, frame_stride=2))),
    )

    # full-size checksums
    sums = {}
    for tag, (w, h, n, first) in {
        "config1_512x512_f0": (512, 512, 1, 0),
        "512x512_acc3": (512, 512, 3, 0),
        "config2_1920x1080_f0": (1920, 1080, 1, 0),
        "config4_3840x2160_f1": (3840, 2160, 1, 1),  # seeds collide above 1920x1080
        "config2_1920x1080_64spp": (1920, 1080, 64, 0),  # the headline frame at full size (about 20 s on 8 cores)
    }.items():
        u = ref.scene(w, h)
        layers = ref.render_frames(u, w, h, n, frame_first=first)
        sums[tag] = {"width": w, "height": h, "frames": n, "frame_first": first,
                     **{k: sha(a) for k, a in zip(LAYERS, layers)},
                     "beauty_mean": float(layers[0][..., :3].mean(dtype=np.float64))}
    (HERE / "ref_checksums.json").write_text(json.dumps(sums, indent=1) + "\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
