"""Host-side plumbing of the multi-GPU path: genomic-range shard planning and the merge of per-shard
sparse edge tables.  The device does the merge-reduce (sqg_merge_edge_tables); `merge_edge_tables` is the same
operation on host arrays, used where the tables have already been gathered to the host and by the CPU (gloo) tests."""
from __future__ import annotations

import numpy as np


def pack_edge_keys(ind1, ind2, head1, head2) -> np.ndarray:
    """(Ind1, Ind2, Head1, Head2) -> the 64-bit sort key of the edge kernel (Edge_t::operator<, src/BPEdge.h:59-70)."""
    return (np.asarray(ind1, np.uint64) << np.uint64(33)) | (np.asarray(ind2, np.uint64) << np.uint64(2)) | \
           (np.asarray(head1, np.uint64) << np.uint64(1)) | np.asarray(head2, np.uint64)


def unpack_edge_keys(keys):
    k = np.asarray(keys, np.uint64)
    return ((k >> np.uint64(33)).astype(np.int32), ((k >> np.uint64(2)) & np.uint64(0x7FFFFFFF)).astype(np.int32),
            ((k >> np.uint64(1)) & np.uint64(1)).astype(bool), (k & np.uint64(1)).astype(bool))


def merge_edge_tables(tables):
    """[(keys uint64, weights int32), ...] -> (unique sorted keys, summed weights): BuildEdges' sort + run-length sum
    (SegmentGraph.cpp:1943-1949) applied to already-reduced per-shard tables."""
    keys = np.concatenate([np.asarray(k, np.uint64) for k, _ in tables]) if tables else np.zeros(0, np.uint64)
    w = np.concatenate([np.asarray(x, np.int64) for _, x in tables]) if tables else np.zeros(0, np.int64)
    if keys.size == 0:
        return keys, w.astype(np.int32)
    order = np.argsort(keys, kind="stable")
    keys, w = keys[order], w[order]
    head = np.r_[True, keys[1:] != keys[:-1]]
    idx = np.flatnonzero(head)
    return keys[idx], np.add.reduceat(w, idx).astype(np.int32)


def plan_shards(ref_id: np.ndarray, world: int):
    """Contiguous record ranges [lo,hi) per rank, cut at chromosome boundaries (never inside a chromosome) and balanced
    by record count.  Returns a list of (lo, hi)."""
    n = int(ref_id.shape[0])
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (world - 1)
    mapped = ref_id[ref_id >= 0]
    cuts = np.flatnonzero(np.r_[True, mapped[1:] != mapped[:-1]])  # first record of every chromosome
    cuts = np.r_[cuts, n]
    bounds = [0]
    for k in range(1, world):
        target = n * k // world
        j = int(np.argmin(np.abs(cuts - target)))
        bounds.append(int(max(cuts[j], bounds[-1])))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]
