"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for the collectives.

The hot path shards with no data-path exchange: Markov chains (and ExactSummation basis ranges) are independent, so
rank r of G owns the chains / indices [r*n/G, (r+1)*n/G) — Philox streams are keyed by the GLOBAL chain id, so the
sampled configurations do not depend on G.  The only communication is the sum of the packed partial sums
{sum w E, sum w |E|^2, sum w, sum w O_k, sum w E O_k*} (one all-reduce per functional), the S-matrix partial
(dense SR) or one P-vector per CG iteration (matrix-free SR).  libangpu owns an NCCL communicator for these sums
(`angpu_comm_init`, csrc/comm.cu): `init_from_env` creates the unique id on rank 0, broadcasts its 128 bytes through
`torch.distributed` and initialises the library's communicator on every rank, so the collectives are plain
`ncclAllReduce` calls on the library's stream -- no Python in the loop.  `ANGPU_COMM=hook` selects the older transport,
a callback into `allreduce_hook` (dist.all_reduce on a zero-copy view of the device pointer).

The reference has no counterpart (single device, no collectives — SURVEY.md §2).
"""
import datetime
import os

import torch
import torch.distributed as dist

from . import api


_stream = None


class _DevArray:
    """Zero-copy view of `count` float64 at a raw CUDA device pointer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def shard_range(total, rank, world):
    """The C ABI's partition (Ensemble::shard in csrc/vmc.hpp): [total*rank//world, total*(rank+1)//world)."""
    begin = total * rank // world
    return begin, total * (rank + 1) // world - begin


_views = {}


def allreduce_hook(ptr, count):
    # the library reduces a handful of grow-only buffers over and over (packed sums, one P-vector per CG product): the
    # zero-copy tensor views are cached by (pointer, count) so that a call costs one dict lookup + dist.all_reduce
    t = _views.get((ptr, count))
    if t is None:
        if len(_views) >= 64:
            _views.clear()
        t = torch.as_tensor(_DevArray(ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
        _views[(ptr, count)] = t
    dist.all_reduce(t, op=dist.ReduceOp.SUM)


def init_from_env(backend="nccl"):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world). Single process: (0, 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    api.setDevice(local_rank)
    # Run the library on a dedicated torch stream made current: NCCL collectives issued by the hook and torch.cuda.Event
    # timing are then ordered with the library's kernels.  (torch's default stream has handle 0, which the C ABI reads
    # as "use your own stream" — never pass it.)
    global _stream
    _stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(_stream)
    api.set_stream(_stream.cuda_stream)
    if world > 1:
        hook = os.environ.get("ANGPU_COMM", "nccl") == "hook"
        if not dist.is_initialized():
            # With the communicator inside the library, torch.distributed only bootstraps (unique id, barriers, the bench's MAX over
            # ranks): its NCCL communicator is NOT created eagerly (no device_id) and the CPU side runs on gloo -- a second NCCL
            # communicator in the process competes with the library's for the NVLS (NVLink SHARP) resources (measured at 8 GPUs:
            # the 524 KB all-reduce of eval_F took 0.10 ms longer with both communicators alive).
            kw = dict(device_id=torch.device("cuda", local_rank)) if (hook and backend == "nccl") else {}
            if os.environ.get("MASTER_ADDR", "127.0.0.1") in ("127.0.0.1", "localhost"):
                os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")       # single node: gloo must not depend on the hostname resolving
            be = backend if (hook or backend != "nccl") else "cpu:gloo,cuda:nccl"
            dist.init_process_group(backend=be, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=int(os.environ.get("ANGPU_NCCL_TIMEOUT_S", "120"))), **kw)
        if hook:
            api.set_allreduce(allreduce_hook)
            # ensembles do not inherit a shard from the callback transport: callers use .set_shard(rank, world)
        else:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8).clone()
            dist.broadcast(uid, src=0)                                  # CPU tensor: gloo
            api.comm_init(bytes(uid.numpy().tobytes()), rank, world)
    return rank, world


def barrier():
    """Host-side barrier over the ranks (a CPU all-reduce: does not create a torch NCCL communicator)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if "gloo" in str(dist.get_backend()):
            dist.all_reduce(torch.zeros(1))
        else:
            dist.barrier()


def max_over_ranks(value):
    """MAX of a host scalar over the ranks."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    if "gloo" in str(dist.get_backend()):
        t = torch.tensor([float(value)], dtype=torch.float64)
    else:
        t = torch.tensor([float(value)], dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shutdown():
    _views.clear()
    api.set_allreduce(None)
    api.synchronize()
    api.comm_destroy()
    api.set_stream(None)
    if dist.is_initialized():
        dist.destroy_process_group()
