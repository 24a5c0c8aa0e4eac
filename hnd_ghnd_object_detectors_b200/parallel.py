"""Data-parallel glue for the student's flat gradient buffer.

The reference wraps the student in DistributedDataParallel (src/mimic_runner.py:141-143): a bucketed
NCCL all-reduce of 25 gradients plus a per-forward buffer broadcast.  Here all trainable tensors live
in ONE flat fp32 buffer (engine.FlatParams), so the exchange is a single all-reduce (2.35 MB) over
NVLink/NVSwitch, followed by the fused Adam kernel that applies 1/world_size.  BN running statistics
are per-rank local (no SyncBN in the reference either); rank 0's are the ones that get saved."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_flat_grad(flat, async_op=False):
    """SUM all-reduce of FlatParams.grad (the averaging is folded into FusedAdam's grad_scale)."""
    if world_size() == 1:
        return None
    return dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM, async_op=async_op)


def broadcast_flat_params(flat, src=0):
    """Make every rank start from rank `src`'s student parameters (DDP does this at construction)."""
    if world_size() > 1:
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
