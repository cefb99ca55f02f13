"""Fabricates dataset files with the exact on-disk layouts (create_binary_db.py:44-87, data/BinaryDbReaderSTB.py:99-121) from
random images / key-points: the real RHD / STB sets are not available offline, so the example drivers only demonstrate the
plumbing unless --db points at a real file."""
import os
import tempfile

import numpy as np


def fake_rhd(n, seed=0):
    rng = np.random.default_rng(seed)
    out = bytearray()
    for _ in range(n):
        out += (rng.normal(scale=0.08, size=(42, 3)) + np.array([0, 0, 0.5])).astype(np.float32).tobytes()
        out += rng.uniform(40, 280, size=(42, 2)).astype(np.float32).tobytes()
        out += np.array([[283.1, 0, 160.0], [0, 283.1, 160.0], [0, 0, 1]], np.float32).tobytes() + b"\x00\x00"
        out += rng.integers(0, 256, size=(320, 320, 3), dtype=np.uint8).tobytes()
        out += rng.integers(0, 34, size=(320, 320), dtype=np.uint8).tobytes()
        out += (rng.uniform(size=42) > 0.2).astype(np.uint8).tobytes()
    return bytes(out)


def fake_stb(n, seed=0):
    rng = np.random.default_rng(seed)
    out = bytearray()
    for _ in range(n):
        out += (rng.normal(scale=60.0, size=(21, 3)) + np.array([0, 0, 600.0])).astype(np.float32).tobytes()
        out += np.concatenate([rng.uniform(30, 600, size=(21, 1)), rng.uniform(30, 460, size=(21, 1)),
                               (rng.uniform(size=(21, 1)) > 0.2).astype(np.float64)], 1).astype(np.float32).tobytes()
        out += rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8).tobytes()
    return bytes(out)


def db_path(arg, kind, n):
    """--db path, or a temporary synthetic file of n records."""
    if arg:
        return arg, None
    tmp = tempfile.NamedTemporaryFile(suffix=".bin", delete=False)
    tmp.write(fake_rhd(n) if kind == "rhd" else fake_stb(n))
    tmp.close()
    return tmp.name, tmp.name


def cleanup(path):
    if path and os.path.exists(path):
        os.unlink(path)
