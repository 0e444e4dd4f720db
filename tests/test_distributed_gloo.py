"""World-size-2 test of the sharded macro <-> micro exchange on CPU (gloo): each rank assembles and
solves its block of micro systems (product kernel bodies through the CPU harness), the coupling
integrals are allgathered with the padded-block layout libmsgpu uses over NCCL, rank 0 'solves' the
macro problem and broadcasts.  The result must equal the single-rank run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import BIO, INPUT, MESH_SUBDIVIDED, ROOT, Emul, configure_bio
from dealiiscale_b200.capi import partition
from dealiiscale_b200.parallel import MacroSharding
from dealiiscale_b200.prm import BioData

M = 3  # 3x3 macro points: an odd total, so the last rank's block is padded


def _points():
    g = np.linspace(-1, 1, M)
    return np.array([(a, b) for b in g for a in g])


def _cycle(xy, u, w):
    e = Emul(BIO, MESH_SUBDIVIDED, 4, 12)
    configure_bio(e, BioData(os.path.join(INPUT, "biocase-curved.prm")))
    e.setup(xy)
    assert e.assemble(u, w) == 0
    rc, _ = e.solve(1e-12, 10000)
    assert rc == 0
    cu, cw = e.coupling()
    e.close()
    return cu, cw


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xy = _points()
    sh = MacroSharding(len(xy), rank, world)
    # the macro iterate lives on rank 0 and is broadcast (msgpu_broadcast_macro)
    u = xy[:, 1] * np.sin(xy[:, 0]) if rank == 0 else np.zeros(len(xy))
    w = xy[:, 0] * np.cos(xy[:, 1]) if rank == 0 else np.zeros(len(xy))
    u, w = sh.broadcast(u), sh.broadcast(w)
    cu, cw = _cycle(sh.local(xy), sh.local(u), sh.local(w))
    cu_all, cw_all = sh.allgather(cu), sh.allgather(cw)
    if rank == 0:
        np.save(out, np.stack([cu_all, cw_all]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = os.path.join(tmp_path, "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    xy = _points()
    cu, cw = _cycle(xy, xy[:, 1] * np.sin(xy[:, 0]), xy[:, 0] * np.cos(xy[:, 1]))
    assert np.array_equal(got[0], cu) and np.array_equal(got[1], cw)


def test_partition_covers_everything_once():
    for n_total, world in ((9, 2), (4225, 8), (5, 8), (33800, 8)):
        seen = []
        for r in range(world):
            k0, k1 = partition(n_total, world, r)
            seen += list(range(k0, k1))
            sh = MacroSharding(n_total, r, world)
            assert (sh.k0, sh.k1) == (k0, k1) and sh.n_local <= sh.block
        assert seen == list(range(n_total))


def test_cpp_driver_rendezvous_between_processes(tmp_path):
    """host/distributed.h: the file-based exchange the multi-process C++ drivers use for the NCCL unique id and the
    peer-memory handles (allgather, broadcast, barrier, clean-up), three processes, no GPU involved."""
    import subprocess
    import sys

    code = ("import ctypes, os, sys; "
            f"lib = ctypes.CDLL({os.path.join(ROOT, 'dealiiscale_b200', '_build', 'libmsdrivers.so')!r}); "
            "sys.exit(0 if lib.msdrv_rendezvous_selftest(5) == int(os.environ['WORLD_SIZE']) else 1)")
    rdzv = os.path.join(tmp_path, "rdzv")
    procs = [subprocess.Popen([sys.executable, "-c", code],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3", LOCAL_RANK=str(r),
                                       MSGPU_RENDEZVOUS_DIR=rdzv)) for r in range(3)]
    assert [p.wait(timeout=120) for p in procs] == [0, 0, 0]
    assert not os.path.exists(rdzv)  # rank 0 removed the exchange directory after the last barrier


def test_coarse_levels_use_only_as_many_ranks_as_have_macro_points():
    """host/distributed.h: with B = ceil(N / R) the last blocks of the partition can be empty (N = 9 on 4 ranks);
    the drivers then run the level on the largest R that leaves no rank empty and the others sit it out."""
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "dealiiscale_b200", "_build", "libmsdrivers.so"))
    cases = ((9, 4), (9, 8), (25, 8), (4225, 8), (1, 8), (9, 2), (81, 8))
    assert [lib.msdrv_usable_ranks(n, w) for n, w in cases] == [3, 5, 7, 8, 1, 2, 8]
    for n in range(1, 200):
        for w in range(1, 9):
            r = lib.msdrv_usable_ranks(n, w)
            blocks = [partition(n, r, k) for k in range(r)]
            assert 1 <= r <= w and all(k1 > k0 for k0, k1 in blocks) and blocks[-1][1] == n
            if r < w:  # one more rank would leave somebody without work
                assert any(k1 <= k0 for k0, k1 in (partition(n, r + 1, k) for k in range(r + 1)))


def _check_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11  # odd: the last block is padded
    sh = MacroSharding(n, rank, world)
    truth = [np.arange(n) * 1.5 + 0.25, -np.arange(n) ** 2 * 1.0]
    good = sh.check_exchange(*truth, local=[sh.local(t) for t in truth])
    broken = [t.copy() for t in truth]
    if rank == 1:
        broken[1][3] += 1e-9  # one value delivered wrong to one rank
    bad = sh.check_exchange(*broken, local=[sh.local(t) for t in truth])
    if rank == 0:
        np.save(out, np.array([good, bad]))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_check_counts_wrong_deliveries(tmp_path):
    """bench.py's `exchange_check` (MacroSharding.check_exchange): what every rank holds after the exchange against
    an independent allgather of the owners' values."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = os.path.join(tmp_path, "check.npy")
    mp.spawn(_check_worker, args=(2, port, out), nprocs=2, join=True)
    good, bad = np.load(out)
    assert good == 0 and bad == 1
