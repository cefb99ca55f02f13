"""On-device decode of the dataset readers' on-disk record formats.

The reference decodes fixed-length records inside TF queue runners (data/BinaryDbReader.py:103-208 for RHD,
data/BinaryDbReaderSTB.py:99-185 for STB).  At thousands of images per second the uint8 -> `x/255 - 0.5` conversion,
the optional 2x sub-sampling (eval_full.py:50) and the host->device copy become the next bottleneck, so the raw records
are copied to the GPU as bytes (4x fewer bytes than fp32 images) and decoded there by one kernel.  Only the raw items
are produced (the readers' derived items -- palm coordinates, dominant-hand selection, crops, score-map targets --
belong to the training / evaluation data pipeline, which is out of scope).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import runtime

RHD_RECORD_BYTES = 2 + 4 * (42 * 3 + 42 * 2 + 9) + 320 * 320 * 3 + 320 * 320 + 42      # 410 520 (data/BinaryDbReader.py:103-124)
STB_RECORD_BYTES = 4 * (21 * 3 + 21 * 3) + 480 * 640 * 3                                # 922 104


def _records(records, record_bytes):
    if isinstance(records, (bytes, bytearray, memoryview)):
        records = np.frombuffer(records, dtype=np.uint8)
    if isinstance(records, np.ndarray):
        records = torch.from_numpy(np.array(records, dtype=np.uint8, copy=True))     # own, writable host copy
    records = records.reshape(-1, record_bytes)
    if not records.is_cuda:
        records = records.pin_memory().to(runtime.default_context().device, non_blocking=True)
    return records


def decode_rhd_records(records):
    """records: bytes / uint8 array / tensor holding B records of 410 520 bytes.
    Returns the raw items of BinaryDbReader.get() as CUDA tensors: image [B,320,320,3] fp32 (u8/255-0.5),
    keypoint_xyz [B,42,3], keypoint_uv [B,42,2] (cast to int32 and back, as the reader does), cam_mat [B,3,3],
    hand_parts [B,320,320] int32, hand_mask [B,320,320,2] int32 (bg, hand = parts > 1), keypoint_vis [B,42] bool."""
    r = runtime.default_context().decode_records(_records(records, RHD_RECORD_BYTES), "rhd", 1)
    h = r["header"]
    B = h.shape[0]
    parts = r["mask"].to(torch.int32)
    hand = parts > 1
    return {
        "image": r["image"],
        "keypoint_xyz": h[:, :126].reshape(B, 42, 3),
        "keypoint_uv": h[:, 126:210].reshape(B, 42, 2).to(torch.int32).to(torch.float32),     # data/BinaryDbReader.py:151-154
        "cam_mat": h[:, 210:219].reshape(B, 3, 3),
        "hand_parts": parts,
        "hand_mask": torch.stack([~hand, hand], 3).to(torch.int32),
        "keypoint_vis": r["visibility"].to(torch.bool),
    }


def decode_stb_records(records, subsample=2):
    """records: B records of 922 104 bytes.  Returns image [B,480/s,640/s,3] fp32 (subsample=2 reproduces
    eval_full.py:50's tf.image.resize_images(image, (240, 320)), i.e. every 2nd pixel under the TF1 legacy kernel),
    keypoint_xyz [B,21,3] (mm, raw order), keypoint_uv [B,21,2], keypoint_vis [B,21] bool."""
    r = runtime.default_context().decode_records(_records(records, STB_RECORD_BYTES), "stb", subsample)
    h = r["header"]
    B = h.shape[0]
    uvv = h[:, 63:126].reshape(B, 21, 3)
    return {"image": r["image"], "keypoint_xyz": h[:, :63].reshape(B, 21, 3), "keypoint_uv": uvv[:, :, :2].contiguous(),
            "keypoint_vis": uvv[:, :, 2] > 0.5}
