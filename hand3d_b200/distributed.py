"""Multi-GPU plumbing: the pipeline is embarrassingly data parallel (every image is independent,
utils/general.py:250,293 even loop per sample), so ranks shard the batch, replicate the weights and run
with no communication; the only exchange is ONE all-gather of a fixed 432-byte per-image record
(coord3d 21x3 f32 | key-points (row, col) 21x2 i32 | center 2 f32 | scale_crop 1 f32) -- SURVEY.md 8(e).
The large score maps stay sharded on their owning GPU.  Backend-agnostic (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_FLOATS = 63 + 42 + 2 + 1   # 108 x 4 B = 432 B per image


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` images for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(coord3d, keypoints_uv, center, scale_crop):
    """[B,108] records.  CUDA tensors: ONE kernel (h3d_pack_records); CPU tensors (gloo tests, host-side tools): torch.cat."""
    B = coord3d.shape[0]
    if coord3d.is_cuda:
        from . import runtime
        return runtime.default_context(coord3d.device).pack_records(coord3d, keypoints_uv, center, scale_crop)
    uv_bits = keypoints_uv.reshape(B, 42).contiguous().view(torch.float32)    # bit-cast, gathered bitwise
    return torch.cat([coord3d.reshape(B, 63), uv_bits, center.reshape(B, 2), scale_crop.reshape(B, 1)], dim=1).contiguous()


def unpack_records(rec):
    B = rec.shape[0]
    return {
        "keypoint_coord3d": rec[:, :63].reshape(B, 21, 3),
        "keypoints_uv": rec[:, 63:105].contiguous().view(torch.int32).reshape(B, 21, 2),
        "center": rec[:, 105:107],
        "scale_crop": rec[:, 107:108],
    }


def gather_records(rec, group=None):
    """all_gather of equally sized [B_local, 108] record tensors -> [world * B_local, 108], rank-major."""
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out


def gather_ragged_records(rec, total: int, group=None):
    """Same for shard_range() shards of unequal size: pads to the largest shard, gathers, strips the padding."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    pad[: sizes[rank]] = rec
    full = gather_records(pad, group).reshape(world, mx, rec.shape[1])
    return torch.cat([full[r, : sizes[r]] for r in range(world)], dim=0)


class P2PGather:
    """Fused record pack + all-gather over NVLink peer memory (h3d_gather_records_p2p): the per-image records are written
    by ONE kernel straight into every peer's symmetric gather buffer (peer-mapped stores, or a multimem store on the NVSwitch
    multicast address), with flag-based completion -- no NCCL launch and no separate pack kernel on the critical path.
    torch.distributed._symmetric_memory only provides the allocation / address exchange (plumbing)."""

    def __init__(self, ctx, max_batch: int, group=None, use_multicast: bool = True):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.ctx, self.C = ctx, C
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.max_batch = max_batch
        self.stride = self.world * max_batch * RECORD_FLOATS                       # floats per parity
        self.buf = symm_mem.empty(2 * self.stride, dtype=torch.float32, device=ctx.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        # dedicated signal words (one uint32 per peer): NOT the handle's signal pad, which hdl.barrier() uses
        self.sig = symm_mem.empty(max(64, self.world), dtype=torch.int32, device=ctx.device)
        self.sig.zero_()
        self.sig_hdl = symm_mem.rendezvous(self.sig, self.group)
        self.mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        self.epoch = 0
        torch.cuda.synchronize(ctx.device)
        self.hdl.barrier()

    def gather(self, coord3d, keypoints_uv, center, scale_crop):
        """-> [world * B, 108] records of all ranks (rank-major): a view of the local gather buffer when B == max_batch, else
        the per-rank slots' first B rows (every rank must then pass the same B; ragged shards read `slots()` instead)."""
        from . import _lib
        B = coord3d.shape[0]
        if B > self.max_batch:
            raise ValueError("batch %d exceeds the gather buffer (%d)" % (B, self.max_batch))
        self.epoch += 1
        C = self.C
        cs = torch.cuda.current_stream().cuda_stream
        _lib.check(self.ctx.lib.h3d_gather_records_p2p(
            self.ctx.h, C.c_void_p(coord3d.data_ptr()), C.c_void_p(keypoints_uv.data_ptr()), C.c_void_p(center.data_ptr()),
            C.c_void_p(scale_crop.data_ptr()), B, self.max_batch, C.c_void_p(int(self.hdl.buffer_ptrs_dev)),
            C.c_void_p(int(self.sig_hdl.buffer_ptrs_dev)), C.c_uint64(self.mc), self.rank, self.world, C.c_uint32(self.epoch),
            C.c_int64(self.stride), C.c_void_p(cs)), "h3d_gather_records_p2p")
        s = self.slots()
        if B == self.max_batch:
            return s.view(self.world * B, RECORD_FLOATS)
        return s[:, :B].reshape(self.world * B, RECORD_FLOATS)

    def slots(self):
        """[world, max_batch, 108] view of the current parity of the local gather buffer (rank r's records in slot r)."""
        off = (self.epoch & 1) * self.stride
        return self.buf[off: off + self.stride].view(self.world, self.max_batch, RECORD_FLOATS)
