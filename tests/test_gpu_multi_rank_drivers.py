"""The C++ drivers with the macro points sharded over one process per GPU (host/distributed.h):
`torchrun --no-python --nproc-per-node 2 solve_*` must print the same convergence tables as the single-process
run — every rank takes the same stop decisions from allgathered per-grid scalars and bit-reproducible macro
solves.  With fewer GPUs than ranks the ranks SHARE device 0 (MSGPU_RANKS_SHARE_DEVICE=1, as tests/test_gpu_p2p.py
does): NCCL refuses two ranks on one device, so that run has no communicator — the coupling integrals travel through
peer memory (the fused kernel, the default path) and the per-grid scalars of the stop test through the host exchange;
the NCCL variants (ncclAllGather, ncclBroadcast) need one GPU per rank (scripts/multi_gpu_drivers.sh, `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

from helpers import INPUT, ROOT

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "dealiiscale_b200", "_build")


def _n_gpus():
    import torch

    return torch.cuda.device_count()


def _run(tmp, name, exe_args, ranks, port, env_extra=None):
    work = os.path.join(tmp, name)
    os.makedirs(os.path.join(work, "results"))
    os.symlink(INPUT, os.path.join(work, "input"))
    exe = os.path.join(BUILD, exe_args[0])
    if ranks == 1:
        cmd = [exe] + exe_args[1:]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", f"--nproc-per-node={ranks}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), exe] + exe_args[1:]
    r = subprocess.run(cmd, cwd=work, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env_extra or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return work, r.stdout


RANKS = int(os.environ.get("MSGPU_TEST_RANKS", "2"))  # 4: coarse levels leave ranks without macro points (they sit out)
QUICK = os.environ.get("MSGPU_TEST_QUICK") == "1"       # fused exchange and broadcast variants only

CASES = [
    ("bio", ["solve_biomath", "input/linear.prm", "3", "4", "0"], "out.txt"),
    ("bio_nosol", ["solve_biomath", "input/biocase.prm", "2", "3", "0"], None),  # residual-based stop, MSGPU_CYCLE_LIMIT
    ("ell", ["solve_elliptic", "input/elliptic_non_linear.prm"], "elliptic_non_linear_convergence_table.txt"),
    ("par", ["solve_parabolic", "input/linear_time.prm", "2"], "linear_time_convergence_table.txt"),
    ("bio_tiny", ["solve_biomath", "input/linear.prm", "1", "3", "0"], "out.txt"),  # 9 macro points
]


@pytest.mark.parametrize("name,args,table", CASES)
def test_several_ranks_reproduce_the_single_process_run(tmp_path, name, args, table):
    shared = _n_gpus() < RANKS
    if shared and name == "bio_tiny" and RANKS > 2:
        pytest.skip("ranks that share a device must all take part in every level")
    base = {"MSGPU_CYCLE_LIMIT": "4"} if name == "bio_nosol" else {}
    one, out1 = _run(str(tmp_path), name + "_1", args, 1, 0, base)
    variants = [("p2p", {}), ("nccl", {"MSGPU_NCCL_ALLGATHER": "1"}), ("hostmacro", {"MSGPU_HOST_MACRO": "1"}),
                ("bcast", {"MSGPU_MACRO_BROADCAST": "1"}),  # rank 0 solves the macro system, ncclBroadcast
                ("bcasthost", {"MSGPU_MACRO_BROADCAST": "1", "MSGPU_HOST_MACRO": "1"}),
                ("p2pfail", {"MSGPU_TEST_FAIL_P2P": "1"})]  # rank 1 cannot export: everybody falls back to NCCL
    if QUICK:
        variants = [variants[0], variants[3], variants[5]]
    if shared:  # no NCCL communicator between ranks on one device: the peer-memory paths only
        variants = [("p2p", {"MSGPU_RANKS_SHARE_DEVICE": "1"}),
                    ("hostmacro", {"MSGPU_RANKS_SHARE_DEVICE": "1", "MSGPU_HOST_MACRO": "1"})]
    for tag, env in variants:
        two, out2 = _run(str(tmp_path), f"{name}_{RANKS}_{tag}", args, RANKS, 29700 + len(tag), dict(base, **env))
        if table:
            t1 = open(os.path.join(one, "results", table)).read()
            t2 = open(os.path.join(two, "results", table)).read()
            assert t1 == t2 and len(t1.split()) > 7, (tag, t1, t2)
        # the printed residual / error lines decide the stop: identical text on rank 0
        pick = lambda s: [ln for ln in s.splitlines() if "residual" in ln or "error" in ln]  # noqa: E731
        assert pick(out1) == pick(out2), tag
        # every rank wrote the micro files of the grids it owns
        if name.startswith("bio"):
            f1 = sorted(os.listdir(os.path.join(one, "results", "micro-solutions")))
            f2 = sorted(os.listdir(os.path.join(two, "results", "micro-solutions")))
            assert f1 == f2 and len(f1) > 2
