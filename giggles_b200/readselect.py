"""Host-side mirror of the reference's ``readselection`` (giggles/readselect.pyx:231-271).

Same name, arguments, return type and error as the reference function that ``giggles/utils.py:71-76`` calls before
every ``GenotypeHMM``; the work is done by ``gg_readselection`` in libgiggles_b200.so (no Python fallback).

``readset`` is anything that iterates over reads which in turn iterate over variants with ``.position`` and
``.quality`` and carry ``.source_id`` — the reference's own ``giggles.core.ReadSet`` qualifies, and so does a list of
:class:`FlatRead`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class FlatVariant:
    position: int
    quality: int = 0


@dataclass
class FlatRead:
    source_id: int = 0
    variants: list = field(default_factory=list)

    def __iter__(self):
        return iter(self.variants)

    def __len__(self):
        return len(self.variants)


def flatten(readset):
    """(read_variant_offset, variant_position, variant_quality, read_source_id) of a read container."""
    off, pos, qual, src = [0], [], [], []
    for read in readset:
        for v in read:
            pos.append(int(v.position))
            qual.append(int(v.quality))
        off.append(len(pos))
        src.append(int(read.source_id))
    return (np.asarray(off, dtype=np.uint64), np.asarray(pos, dtype=np.int32), np.asarray(qual, dtype=np.int32),
            np.asarray(src, dtype=np.int32))


def readselection(readset, max_cov, preferred_source_ids=None, bridging=True):
    """Indices of the reads to keep so that no variant is covered more than ``max_cov`` times.

    Returns a ``set`` like the reference; raises ``ValueError`` if a read covers fewer than two variants
    (readselect.pyx:249-252)."""
    off, pos, qual, src = flatten(readset)
    sel = capi.readselection_flat(off, pos, qual, src, max_cov, preferred_source_ids, bridging)
    return set(int(i) for i in sel)
