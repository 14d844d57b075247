"""The N > 1 path on CPU: world_size-2 `gloo` run of the sharding + packed all-reduce + finalise logic, with the CPU
oracle standing in for the kernels (chains are keyed by GLOBAL id, so shards reproduce the single-process run)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

CHAINS, PER_CHAIN, SWEEPS, THERM, SEED = 12, 2, 2, 3, 4321


def packed_partials(port, psi, op, chain0, nchains, global_samples):
    """Layout of TDVP::packed (csrc/vmc.hpp): [sum wE (2), sum w|E|^2, sum w, sum w O_k (2P), sum w E conj O_k (2P)]."""
    mc = port.MonteCarlo(PER_CHAIN * nchains, SWEEPS, THERM, nchains, seed=SEED, chain0=chain0)
    confs, _ = mc.sample(psi)
    _, el, O = port.eval_samples(psi, op, confs, want_O=True)
    w = np.full(len(confs), 1.0 / global_samples)
    head = np.array([np.sum(w * el), np.sum(w * np.abs(el) ** 2) + 1j * np.sum(w)])
    return np.concatenate([head, w @ O, (w * el) @ np.conj(O)]), confs


def finalise(packed, P):
    E = packed[0]
    return E, packed[2 + P:2 + 2 * P] - E * np.conj(packed[2:2 + P])


def _worker(rank, world, port_file, out_file):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = open(port_file).read().strip()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from annongpu_b200.distributed import shard_range
    from annongpu_b200 import factories as F
    from oracle import port_oracle as P
    from helpers import make_op, make_psi, zoo
    spec, H, N = zoo()["rbm10"]
    psi, op = make_psi(P, spec), make_op(P, H)
    c0, cn = shard_range(CHAINS, rank, world)
    packed, confs = packed_partials(P, psi, op, c0, cn, CHAINS * PER_CHAIN)
    t = torch.from_numpy(np.ascontiguousarray(packed).view(np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                       # the hook's job (annongpu_b200/distributed.py)
    gathered = [None] * world
    dist.all_gather_object(gathered, (c0, cn, confs))
    # the bench's host-side helpers on the gloo side of the process group (no torch NCCL communicator is created for them)
    from annongpu_b200 import distributed as D
    D.barrier()
    assert D.max_over_ranks(10.0 + rank) == 10.0 + (world - 1)
    if rank == 0:
        np.savez(out_file, packed=t.numpy().view(np.complex128), shards=np.array([(g[0], g[1]) for g in gathered]),
                 confs=np.concatenate([g[2].reshape(PER_CHAIN, g[1], -1) for g in gathered], axis=1).reshape(CHAINS * PER_CHAIN, -1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process(tmp_path, port):
    from helpers import make_op, make_psi, zoo
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_file, out_file = tmp_path / "port", tmp_path / "out.npz"
    port_file.write_text(str(s.getsockname()[1]))
    s.close()
    mp.spawn(_worker, args=(2, str(port_file), str(out_file)), nprocs=2, join=True)
    got = np.load(out_file)
    spec, H, N = zoo()["rbm10"]
    psi, op = make_psi(port, spec), make_op(port, H)
    P = psi.num_params
    ref_packed, ref_confs = packed_partials(port, psi, op, 0, CHAINS, CHAINS * PER_CHAIN)
    assert [tuple(x) for x in got["shards"]] == [(0, 6), (6, 6)]
    assert np.array_equal(got["confs"], ref_confs)                  # same chains whatever the world size
    assert np.allclose(got["packed"], ref_packed, rtol=1e-12, atol=1e-14)
    E, Fv = finalise(got["packed"], P)
    E0, F0 = finalise(ref_packed, P)
    assert abs(E - E0) <= 1e-12 * abs(E0) and np.allclose(Fv, F0, rtol=1e-10, atol=1e-13)


def test_shard_range_partitions_exactly():
    from annongpu_b200.distributed import shard_range
    for total in (1, 7, 8192, 65536, 131072):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == total
            for (b0, n0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + n0 == b1
