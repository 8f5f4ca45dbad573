"""Multi-GPU driver logic: chromosome-level sharding of independent chains.

The reference builds one GenotypeHMM per chromosome (giggles/cli/genotype.py:196-221,313) and the
chains share nothing, so the only multi-GPU structure is: assign chains to ranks, run, gather the
posteriors on rank 0.  One process per GPU; `torch.distributed` is plumbing only (the final
gather), there is no collective on the data path.  Without a process group (``dist=None``) and with
``world_size > 1`` every rank writes its share to ``result_dir`` instead (one .npz per rank) and
:func:`collect_result_dir` assembles them — the torch-free form of the same gather.
"""
from __future__ import annotations

import os
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


def run_sharded(problems, run_one: Callable, rank: int, world_size: int, dist=None, device=None, assign=None,
                result_dir: str | None = None):
    """Run ``run_one(problem) -> flat float64 posteriors`` on this rank's share of ``problems`` and gather
    everything on rank 0 (returns the list of per-chain arrays there, None elsewhere).

    ``dist`` is ``torch.distributed`` (already initialised) or None.  With the NCCL backend the transfers run
    on ``device``; with gloo on the CPU.  Shares are uneven (chromosomes are): every rank sends exactly its
    own posteriors, rank 0 posts one receive per peer — no padding to the largest share.
    """
    if assign is None:
        assign = lpt_assign([chain_cost(p) for p in problems], world_size)
    mine = {i: np.asarray(run_one(problems[i]), dtype=np.float64) for i in assign[rank]}
    if world_size == 1:
        return [mine[i] for i in range(len(problems))]
    sizes = [int(p.gl_offsets()[-1]) for p in problems]
    if dist is None:
        if result_dir is None:
            raise ValueError("run_sharded: world_size > 1 needs a process group (dist) or a result_dir")
        os.makedirs(result_dir, exist_ok=True)
        tmp = os.path.join(result_dir, f".rank{rank}.tmp.npz")
        np.savez(tmp, **{f"chain{i}": v for i, v in mine.items()})
        os.replace(tmp, os.path.join(result_dir, f"rank{rank}.npz"))
        return None
    import torch
    dev = device if device is not None else "cpu"
    share = [sum(sizes[i] for i in idx) for idx in assign]
    if rank != 0:
        if share[rank]:
            buf = torch.from_numpy(np.concatenate([mine[i] for i in assign[rank]])).to(dev)
            dist.send(buf, dst=0)
        return None
    out = [None] * len(problems)
    for i in assign[0]:
        out[i] = mine[i]
    bufs, reqs = {}, []
    for r in range(1, world_size):
        if share[r]:
            bufs[r] = torch.empty(share[r], dtype=torch.float64, device=dev)
            reqs.append(dist.irecv(bufs[r], src=r))
    for q in reqs:
        q.wait()
    for r, b in bufs.items():
        g = b.cpu().numpy()
        o = 0
        for i in assign[r]:
            out[i] = g[o:o + sizes[i]].copy()
            o += sizes[i]
    return out


def collect_result_dir(problems, world_size: int, result_dir: str, assign=None):
    """Assemble the per-rank files written by :func:`run_sharded` (``dist=None``) into the per-chain list."""
    if assign is None:
        assign = lpt_assign([chain_cost(p) for p in problems], world_size)
    out = [None] * len(problems)
    for r in range(world_size):
        with np.load(os.path.join(result_dir, f"rank{r}.npz")) as d:
            for i in assign[r]:
                out[i] = d[f"chain{i}"]
    return out
