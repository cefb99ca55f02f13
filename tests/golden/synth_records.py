"""Seeded synthetic dataset records in the reference's on-disk formats (create_binary_db.py:44-87 for RHD,
data/BinaryDbReaderSTB.py:99-121 for STB), shared by tests/golden/make_golden_reference_reader.py (which runs the unmodified
reader classes on them) and by the tests (which regenerate the same bytes instead of storing 4 MB of records)."""
import numpy as np

RHD_BYTES = 2 + 4 * (42 * 3 + 42 * 2 + 9) + 320 * 320 * 3 + 320 * 320 + 42      # 410 520
STB_BYTES = 4 * (21 * 3 + 21 * 3) + 480 * 640 * 3                                # 922 104


def rhd_records(n=4, seed=71):
    """n records covering: left / right dominant hand, invisible key-points, no visible key-point at all, key-points outside
    the image, an exact left / right pixel tie (-> right hand, the reference uses 'greater')."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        xyz = (rng.normal(scale=0.08, size=(42, 3)) + np.array([0.0, 0.0, 0.55])).astype(np.float32)
        uv = rng.uniform(40.0, 280.0, size=(42, 2)).astype(np.float32)
        if i % 4 == 3:                                   # some key-points outside the image (clamping of the crop extent)
            uv[5] = (-12.5, 100.25); uv[30] = (340.75, 310.5); uv[9] = (150.0, -3.0)
        K = np.array([[283.1, 0.0, 160.0], [0.0, 283.1, 160.0], [0.0, 0.0, 1.0]], np.float32)
        img = rng.integers(0, 256, size=(320, 320, 3), dtype=np.uint8)
        parts = np.zeros((320, 320), np.uint8)
        parts[40:120, 50:130] = 1                        # person
        la, ra = [(70, 50), (40, 70), (50, 50), (64, 60)][i % 4]
        parts[100:100 + la, 30:30 + la] = rng.integers(2, 18, size=(la, la), dtype=np.uint8)       # left-hand part ids 2..17
        parts[180:180 + ra, 200:200 + ra] = rng.integers(18, 34, size=(ra, ra), dtype=np.uint8)    # right-hand part ids 18..33
        if i % 4 == 2:
            parts[180:230, 200:250] = 20                 # 50x50 right = 50x50 left pixels: tie
        vis = (rng.uniform(size=42) > 0.25).astype(np.uint8)
        if i % 4 == 1:
            vis[:] = 1
        if i % 4 == 2:
            vis[:] = 0                                   # no visible key-point: crop size falls back to the clamp (50)
        rec = xyz.tobytes() + uv.tobytes() + K.tobytes() + b"\x00\x00" + img.tobytes() + parts.tobytes() + vis.tobytes()
        assert len(rec) == RHD_BYTES
        out.append(rec)
    return out


def stb_records(n=2, seed=72):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        xyz = (rng.normal(scale=60.0, size=(21, 3)) + np.array([0.0, 0.0, 600.0])).astype(np.float32)        # millimetres
        uvv = np.concatenate([rng.uniform(30.0, 600.0, size=(21, 2)), (rng.uniform(size=(21, 1)) > 0.2).astype(np.float64)], 1).astype(np.float32)
        uvv[:, 1] = np.minimum(uvv[:, 1], 470.0)
        img = rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
        rec = xyz.tobytes() + uvv.tobytes() + img.tobytes()
        assert len(rec) == STB_BYTES
        out.append(rec)
    return out
