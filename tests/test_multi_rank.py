"""World-size-2 run of the sharding driver on CPU (gloo).  The per-chain compute is the CPU oracle here
(test infrastructure); on the GPU box bench.py plugs the CUDA path into the same driver."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assignment_is_balanced_and_deterministic():
    from giggles_b200.shard import lpt_assign
    # GRCh38 chromosome lengths in Mb (chr1..22, X, Y)
    L = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    for ws in (2, 4, 8):
        a = lpt_assign(L, ws)
        assert sorted(sum(a, [])) == list(range(len(L)))
        loads = [sum(L[i] for i in idx) for idx in a]
        assert max(loads) / (sum(L) / ws) < 1.08
        assert a == lpt_assign(L, ws)
    assert lpt_assign([], 2) == [[], []]
    assert lpt_assign([5.0], 4) == [[0], [], [], []]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from giggles_b200.shard import run_sharded
    from giggles_b200.synth import make_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # uneven shares: chains of very different lengths and allele counts (exact-size send / recv, no padding)
    probs = [make_problem([40, 6, 9, 25, 3][i], 4 + i % 3, coverage=4, max_coverage=4, read_len_mean=1500, seed=70 + i,
                          multiallelic_frac=0.3, max_alleles=4) for i in range(5)]
    out = run_sharded(probs, oracle.oracle_run, rank, world, dist=dist)
    dist.barrier()
    if rank == 0:
        ok = all(np.array_equal(out[i], oracle.oracle_run(probs[i])) for i in range(len(probs)))
        q.put(ok)
    dist.destroy_process_group()


def test_torch_free_gather_through_result_files(tmp_path):
    # dist=None with world_size 2: every rank leaves its share in result_dir, collect_result_dir assembles them
    sys.path.insert(0, ROOT)
    import oracle
    from giggles_b200.shard import collect_result_dir, run_sharded
    from giggles_b200.synth import make_problem
    probs = [make_problem(8 + 2 * i, 4, coverage=3, max_coverage=3, read_len_mean=1500, seed=90 + i) for i in range(5)]
    for r in range(2):
        assert run_sharded(probs, oracle.oracle_run, r, 2, result_dir=str(tmp_path)) is None
    out = collect_result_dir(probs, 2, str(tmp_path))
    assert all(np.array_equal(out[i], oracle.oracle_run(probs[i])) for i in range(5))
    with pytest.raises(ValueError):
        run_sharded(probs, oracle.oracle_run, 0, 2)


def test_two_ranks_gloo_gather():
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
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
