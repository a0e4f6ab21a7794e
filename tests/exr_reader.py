"""A small OpenEXR scanline reader (NO_COMPRESSION / ZIPS / ZIP), for tests only.

Independent check of quasi_b200/host/src/exr_writer.cpp: parses the header,
inflates each chunk, undoes the predictor and byte interleave, and returns
{channel name: 2-D numpy array} with the stored pixel type (float16/float32).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


def read_exr(path):
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    assert magic == 20000630, "not an OpenEXR file"
    assert version & 0xFF == 2 and version >> 8 == 0, "single-part scanline files only"
    pos = 8
    attrs = {}

    def cstr():
        nonlocal pos
        end = data.index(b"\0", pos)
        s = data[pos:end].decode()
        pos = end + 1
        return s

    while data[pos] != 0:
        name, typ = cstr(), cstr()
        (size,) = struct.unpack_from("<i", data, pos)
        pos += 4
        attrs[name] = (typ, data[pos:pos + size])
        pos += size
    pos += 1

    channels = []
    raw = attrs["channels"][1]
    p = 0
    while raw[p] != 0:
        end = raw.index(b"\0", p)
        name = raw[p:end].decode()
        p = end + 1
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", raw, p)
        p += 16
        assert xs == 1 and ys == 1
        channels.append((name, {1: np.float16, 2: np.float32, 0: np.uint32}[ptype]))
    assert [c[0] for c in channels] == sorted(c[0] for c in channels), "channels must be sorted"
    compression = attrs["compression"][1][0]
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    width, height = x1 - x0 + 1, y1 - y0 + 1
    lines_per_chunk = {0: 1, 2: 1, 3: 16}[compression]
    n_chunks = (height + lines_per_chunk - 1) // lines_per_chunk
    offsets = struct.unpack_from(f"<{n_chunks}Q", data, pos)

    out = {name: np.zeros((height, width), dtype=dt) for name, dt in channels}
    line_bytes = sum(np.dtype(dt).itemsize * width for _, dt in channels)
    for off in offsets:
        y, size = struct.unpack_from("<ii", data, off)
        payload = data[off + 8:off + 8 + size]
        lines = min(lines_per_chunk, height - (y - y0))
        expect = line_bytes * lines
        if compression != 0 and size < expect:
            tmp = np.frombuffer(zlib.decompress(payload), dtype=np.uint8).astype(np.int64)
            assert tmp.size == expect
            tmp = np.cumsum(np.concatenate([tmp[:1], tmp[1:] - 128])) & 0xFF  # undo the delta predictor
            tmp = tmp.astype(np.uint8)
            half = (expect + 1) // 2
            buf = np.empty(expect, dtype=np.uint8)
            buf[0::2] = tmp[:half]
            buf[1::2] = tmp[half:]
            payload = buf.tobytes()
        assert len(payload) == expect
        p = 0
        for row in range(lines):
            for name, dt in channels:
                n = np.dtype(dt).itemsize * width
                out[name][y - y0 + row] = np.frombuffer(payload, dtype=dt, count=width, offset=p)
                p += n
    return out, attrs
