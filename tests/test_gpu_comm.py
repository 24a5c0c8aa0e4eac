"""The C-ABI NCCL helper (include/ghnd_b200.h ghnd_comm_*): communicator from a unique id,
in-place SUM all-reduce and broadcast of a flat fp32 buffer.  world_size 1 runs on any GPU box; the
2-rank case needs two GPUs (gpurun --gpus 2) and is skipped otherwise."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_world_size_one():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from hnd_ghnd_object_detectors_b200.parallel import FlatComm
    torch.cuda.set_device(0)
    comm = FlatComm(1, 0)
    x = torch.arange(1000, dtype=torch.float32, device="cuda")
    ref = x.clone()
    comm.allreduce(x)
    comm.broadcast(x, 0)
    torch.cuda.synchronize()
    assert torch.equal(x, ref)
    comm.close()


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "GHND_NCCL_DIRECT": "1"})
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hnd_ghnd_object_detectors_b200 import parallel
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # side channel for the unique id only
    torch.manual_seed(rank)
    flat = FlatParams([("w", torch.nn.Parameter(torch.randn(300, 7, device="cuda"))),
                       ("b", torch.nn.Parameter(torch.randn(5, device="cuda")))])
    parallel.broadcast_flat_params(flat, 0)
    flat.grad.copy_(torch.arange(flat.total, dtype=torch.float32, device="cuda") * (rank + 1))
    parallel.allreduce_flat_grad(flat)
    torch.cuda.synchronize()
    q.put((rank, flat.flat.cpu(), flat.grad.cpu()))
    parallel.FlatComm.get().close()
    dist.destroy_process_group()


def test_comm_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (_, p0, g0), (_, p1, g1) = res
    assert torch.equal(p0, p1)  # rank 0's parameters everywhere
    expect = torch.arange(p0.numel(), dtype=torch.float32) * 3
    assert torch.equal(g0, expect) and torch.equal(g1, expect)
