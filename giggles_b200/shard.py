"""Multi-GPU driver logic: chromosome-level sharding of independent chains.

The reference builds one GenotypeHMM per chromosome (giggles/cli/genotype.py:196-221,313) and the
chains share nothing, so the only multi-GPU structure is: assign chains to ranks, run, gather the
posteriors on rank 0.  One process per GPU; `torch.distributed` is plumbing only (barrier + the
final gather), there is no collective on the data path.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def lpt_assign(costs: Sequence[float], world_size: int) -> list[list[int]]:
    """Longest-processing-time-first: chains sorted by cost (desc), each to the least loaded rank.
    Deterministic (ties by index), so every rank computes the same assignment without talking."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return out


def chain_cost(problem) -> float:
    """State-columns of a chain (sum_c 2^u_c R^2): what the sweep's HBM traffic scales with."""
    return float(problem.state_columns())


def run_sharded(problems, run_one: Callable, rank: int, world_size: int, dist=None, device=None):
    """Run ``run_one(problem) -> flat float64 posteriors`` on this rank's share of ``problems`` and gather
    everything on rank 0 (returns the list of per-chain arrays there, None elsewhere).

    ``dist`` is ``torch.distributed`` (already initialised) or None for a single process.  With the
    NCCL backend the gather runs on ``device``; with gloo on the CPU.
    """
    assign = lpt_assign([chain_cost(p) for p in problems], world_size)
    mine = {i: np.asarray(run_one(problems[i]), dtype=np.float64) for i in assign[rank]}
    if dist is None or world_size == 1:
        return [mine[i] for i in range(len(problems))]
    import torch
    sizes = [int(p.gl_offsets()[-1]) for p in problems]
    dev = device if device is not None else "cpu"
    # one flat buffer per rank, padded to the largest share, then a plain gather to rank 0
    share = [sum(sizes[i] for i in idx) for idx in assign]
    pad = max(max(share), 1)
    buf = torch.zeros(pad, dtype=torch.float64, device=dev)
    o = 0
    for i in assign[rank]:
        buf[o:o + sizes[i]] = torch.from_numpy(mine[i]).to(dev)
        o += sizes[i]
    gathered = [torch.empty_like(buf) for _ in range(world_size)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * len(problems)
    for r, idx in enumerate(assign):
        g = gathered[r].cpu().numpy()
        o = 0
        for i in idx:
            out[i] = g[o:o + sizes[i]].copy()
            o += sizes[i]
    return out
