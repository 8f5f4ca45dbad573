"""configs[2]'s partitioning on the GPU: chains assigned longest-first to two ranks, the CUDA path as the per-chain
compute, uneven shares gathered on rank 0 (giggles_b200/shard.py; reference unit = one GenotypeHMM per chromosome,
giggles/cli/genotype.py:196-221).  With two or more devices the ranks use one GPU each and NCCL carries the gather;
on a one-GPU box both ranks share device 0 and the gather goes through gloo (NCCL refuses two ranks on one device).
Either way the gathered posteriors must be bit-identical to single-process runs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _problems():
    from giggles_b200.synth import make_problem
    # chromosome-like: lengths far apart, one chain with 2^9-block columns, tagged and untagged ones
    return [make_problem([60, 14, 33, 8, 21, 45][i], 88, coverage=8, max_coverage=[5, 9, 1, 6, 3, 4][i], read_len_mean=2500,
                         tags=["none", "none", "all", "mixed", "none", "mixed"][i], seed=1300 + i, window=True) for i in range(6)]


def _worker(rank, world, port, ndev, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from giggles_b200 import capi
    from giggles_b200.shard import run_sharded
    dev_id = rank % ndev
    if ndev >= world:
        torch.cuda.set_device(dev_id)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev_id))
        dev = torch.device("cuda", dev_id)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = None
    probs = _problems()
    out = run_sharded(probs, lambda p: capi.hmm_run(p, device=dev_id).values, rank, world, dist=dist, device=dev)
    dist.barrier()
    if rank == 0:
        q.put([o.tobytes() for o in out])
    else:
        assert out is None
    dist.destroy_process_group()


def test_two_ranks_cuda_path_matches_single_process(have_gpu):
    if not have_gpu:
        pytest.fail("no CUDA device: the GenotypeHMM path has no CPU fallback")
    import torch.multiprocessing as mp
    from giggles_b200 import capi
    ndev = capi.lib().gg_device_count()
    single = [capi.hmm_run(p, device=0).values for p in _problems()]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ndev, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(got) == len(single)
    for a, b in zip(single, got):
        assert a.tobytes() == b
