"""Data-parallel glue for the student's flat gradient buffer.

The reference wraps the student in DistributedDataParallel (src/mimic_runner.py:141-143): a bucketed
NCCL all-reduce of 25 gradients plus a per-forward buffer broadcast.  Here all trainable tensors live
in ONE flat fp32 buffer (engine.FlatParams), so the exchange is a single all-reduce (2.35 MB) over
NVLink/NVSwitch, followed by the fused Adam kernel that applies 1/world_size.  BN running statistics
are per-rank local (no SyncBN in the reference either); rank 0's are the ones that get saved."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatComm(object):
    """The C-ABI NCCL communicator (ghnd_comm_*): rank 0's unique id travels through torch.distributed's
    already-initialised process group (object broadcast), then every rank binds its own device.  Used for
    the flat-gradient all-reduce and the start-up parameter broadcast when GHND_NCCL_DIRECT=1; otherwise
    the same two collectives go through torch.distributed (also NCCL)."""
    _instance = None

    def __init__(self, world, rank):
        uid = (ctypes.c_char * 128)()
        if rank == 0:
            _lib.call("ghnd_comm_unique_id", ctypes.cast(uid, ctypes.c_void_p))
        if world > 1:
            box = [bytes(uid.raw)]
            dist.broadcast_object_list(box, src=0)
            uid = (ctypes.c_char * 128).from_buffer_copy(box[0])
        self._h = ctypes.c_void_p()
        _lib.call("ghnd_comm_init_from_unique_id", ctypes.cast(uid, ctypes.c_void_p), world, rank, ctypes.byref(self._h))
        self.world, self.rank = world, rank

    @classmethod
    def get(cls):
        if cls._instance is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            cls._instance = cls(world_size(), rank)
        return cls._instance

    def allreduce(self, buf):
        _lib.call("ghnd_comm_allreduce_flat", self._h, _lib.ptr(buf), buf.numel(), _lib.stream_ptr())

    def broadcast(self, buf, root=0):
        _lib.call("ghnd_comm_broadcast_flat", self._h, _lib.ptr(buf), buf.numel(), int(root), _lib.stream_ptr())

    def close(self):
        if self._h:
            _lib.load().ghnd_comm_destroy(self._h)
            self._h = ctypes.c_void_p()
        if FlatComm._instance is self:
            FlatComm._instance = None


def _direct():
    return os.environ.get("GHND_NCCL_DIRECT", "0") == "1"


def allreduce_flat_grad(flat, async_op=False):
    """SUM all-reduce of FlatParams.grad (the averaging is folded into FusedAdam's grad_scale)."""
    if world_size() == 1:
        return None
    if _direct() and flat.grad.is_cuda:
        FlatComm.get().allreduce(flat.grad)
        return None
    return dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM, async_op=async_op)


def broadcast_flat_params(flat, src=0):
    """Make every rank start from rank `src`'s student parameters (DDP does this at construction)."""
    if world_size() > 1:
        if _direct() and flat.flat.is_cuda:
            FlatComm.get().broadcast(flat.flat, src)
        else:
            dist.broadcast(flat.flat, src=src)


def broadcast_buffers(module, src=0):
    """BN running statistics etc. of rank `src` to every rank (DDP broadcasts buffers too)."""
    if world_size() > 1:
        for b in module.buffers():
            dist.broadcast(b, src=src)


def shard_indices(n_items, rank, world):
    """DistributedSampler-style partition (data_util.py:28-30): item i goes to rank i % world, padded
    so that every rank gets the same count."""
    per = (n_items + world - 1) // world
    idx = [(rank + k * world) % max(n_items, 1) for k in range(per)]
    return idx
