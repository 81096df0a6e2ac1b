"""
Raw-byte set semantics of the reference's containers, on device-resident rows.

The reference keys Python dicts by `values.tobytes()` (ValueFunction ctor src/mdp.py:668-669, `extend` :773-774,
BeliefSet.union src/pomdp.py:585-606) -- on its GPU path that is one D2H copy per row.  Here rows stay on the
device: `pbvi_row_hash` gives a 128-bit key per row and every key match is confirmed bytewise on the device
(`pbvi_rows_equal` / `pbvi_confirm_groups`), so the result is exact, not probabilistic.  The backup itself groups its
tuples and row keys on the device (`DeviceModel.group_keys`); the helpers below are the host form of the same grouping,
used by the container classes (constructor dedup, `extend`, `union`), by the 'rows' exchange of the sharded backup and as
the reference implementation the device grouping is tested against.
"""
from __future__ import annotations

import numpy as np
import torch

_KEY = np.dtype([('a', '<i8'), ('b', '<i8')])


def _groups_from_sorted(srt: np.ndarray, boundary: np.ndarray):
    """(first, last, inverse) in first-occurrence order from a key-sorted permutation `srt` whose ties keep the original
    order, and the run-start flags `boundary` of the sorted keys."""
    n = srt.shape[0]
    starts = np.flatnonzero(boundary)
    first = srt[starts]
    last = srt[np.append(starts[1:], n) - 1]
    order = np.argsort(first, kind='stable')                 # renumber the groups by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    inverse = np.empty(n, dtype=np.int64)
    inverse[srt] = rank[np.cumsum(boundary) - 1]
    return first[order].astype(np.int64), last[order].astype(np.int64), inverse


def _sort_with_positions(key: np.ndarray, key_bits: int):
    """Stable sort of non-negative int64 keys of at most `key_bits` bits by sorting (key << nb | position) as plain values
    (NumPy's SIMD sort) -- several times faster than a stable argsort.  Returns (positions in key order, sorted keys)."""
    n = key.shape[0]
    nb = max(1, int(np.ceil(np.log2(max(n, 2)))))
    assert key_bits + nb <= 62
    comb = (key << nb) | np.arange(n, dtype=np.int64)
    comb.sort()
    return comb & ((1 << nb) - 1), comb >> nb


def group_by_key(hashes: np.ndarray):
    """
    hashes [n,2] int64 -> (first [g], last [g], inverse [n]) with groups numbered in order of first occurrence --
    the insertion order of a Python dict keyed by the row bytes.
    """
    n = hashes.shape[0]
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    h = np.ascontiguousarray(hashes, dtype=np.int64)
    # sort on a truncated mix of the two halves, then compare the full 128-bit keys inside the sorted order; if two different
    # keys share the truncated mix (so that a run of equal sort keys is not one full key) fall back to a lexicographic sort
    nb = max(1, int(np.ceil(np.log2(max(n, 2)))))
    kb = 62 - nb
    with np.errstate(over='ignore'):
        mix = (h[:, 0] ^ (h[:, 1] * np.int64(-7046029254386353131))) & np.int64((1 << kb) - 1)
    srt, ms = _sort_with_positions(mix, kb)
    hs = h[srt]
    boundary = np.empty(n, dtype=bool)
    boundary[0] = True
    boundary[1:] = (hs[1:, 0] != hs[:-1, 0]) | (hs[1:, 1] != hs[:-1, 1])
    if np.any(boundary[1:] != (ms[1:] != ms[:-1])):
        srt = np.lexsort((np.arange(n), h[:, 1], h[:, 0]))
        hs = h[srt]
        boundary[1:] = (hs[1:, 0] != hs[:-1, 0]) | (hs[1:, 1] != hs[:-1, 1])
    return _groups_from_sorted(srt, boundary)


def unique_rows_first(keys: np.ndarray):
    """Unique rows of an int array in order of first occurrence: (first [u], last [u], inverse [n])."""
    n = keys.shape[0]
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    span = keys.max(axis=0) + 1
    bits = np.ceil(np.log2(np.maximum(span, 2))).astype(np.int64)
    nb = max(1, int(np.ceil(np.log2(max(n, 2)))))
    if int(bits.sum()) + nb <= 62 and keys.min() >= 0:          # pack a row into one int64 (the common case)
        packed = np.zeros(n, dtype=np.int64)
        for j in range(keys.shape[1]):
            packed = (packed << int(bits[j])) | keys[:, j]
        srt, ks = _sort_with_positions(packed, int(bits.sum()))
        boundary = np.empty(n, dtype=bool)
        boundary[0] = True
        boundary[1:] = ks[1:] != ks[:-1]
        return _groups_from_sorted(srt, boundary)
    _, first, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(n)
    order = np.argsort(first, kind='stable')
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    inverse = rank[inverse]
    first = first[order]
    last = np.zeros_like(first)
    last[inverse] = np.arange(n)                             # ascending assignment: the largest position of each group wins
    return first, last, inverse


def dedup_rows(dev, rows: torch.Tensor, hashes: np.ndarray | None = None):
    """
    Dict-insertion dedup of device rows [n,S]: returns (first, last, hashes, inverse) where `first[g]` is the position of the
    first occurrence of group g (groups in first-occurrence order) and `last[g]` of its last occurrence
    (the reference keeps the first POSITION and the last VALUE, i.e. the last action: src/mdp.py:668-669);
    `inverse[i]` is the group of row i.
    """
    n = rows.shape[0]
    if hashes is None:
        hashes = dev.row_hash(rows).cpu().numpy() if n else np.zeros((0, 2), dtype=np.int64)
    first, last, inverse = group_by_key(hashes)
    if first.shape[0] == n:
        return first, last, hashes, inverse
    # confirm every key match bytewise against the group's first row
    dup = np.flatnonzero(first[inverse] != np.arange(n))
    flags = dev.rows_equal(rows, first[inverse[dup]].astype(np.int32), rows, dup.astype(np.int32)).cpu().numpy()
    if not flags.all():
        return _dedup_exact_host(rows, hashes)
    return first, last, hashes, inverse


def _dedup_exact_host(rows: torch.Tensor, hashes: np.ndarray):
    """128-bit key collision between different rows (never observed): redo the grouping on the actual bytes."""
    host = rows.cpu().numpy()
    table = {}
    inverse = np.empty(host.shape[0], dtype=np.int64)
    for i in range(host.shape[0]):
        k = host[i].tobytes()
        if k in table:
            table[k][1] = i
        else:
            table[k] = [i, i, len(table)]
        inverse[i] = table[k][2]
    first = np.array([v[0] for v in table.values()], dtype=np.int64)
    last = np.array([v[1] for v in table.values()], dtype=np.int64)
    return first, last, hashes, inverse
