"""Host-side model of how rows are routed between GPUs.

``partition_ids`` is a numpy restatement of the device routing function of partition.cu
(``partition_of(fmix64(packed key))``) for keys that pack into 64 bits.  It is used
  * by the world_size-2 gloo test, which runs the q8 exchange on CPU ranks with exactly the routing the
    GPUs use, and
  * by the GPU tests, which check flockgpu_hash_partition row by row against it.
``round_robin`` is the batch dealing of RepartitionExec::RoundRobinBatch (planner.rs:91) used to shard
the input stream across ranks: rank r of N owns batches r, r + N, r + 2N, ...
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa

_M1 = np.uint64(0xFF51AFD7ED558CCD)
_M2 = np.uint64(0xC4CEB9FE1A85EC53)


def fmix64(k: np.ndarray) -> np.ndarray:
    """Murmur3 64-bit finaliser (device_utils.cuh: fmix64)."""
    k = k.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= _M1
        k ^= k >> np.uint64(33)
        k *= _M2
        k ^= k >> np.uint64(33)
    return k


def _as_u64(arr: pa.Array) -> tuple[np.ndarray, int]:
    t = arr.type
    if pa.types.is_int32(t) or pa.types.is_uint32(t):
        return arr.to_numpy(zero_copy_only=False).astype(np.int32 if pa.types.is_int32(t) else np.uint32).view(np.uint32).astype(np.uint64), 4
    if pa.types.is_timestamp(t):
        return arr.cast(pa.int64()).to_numpy(zero_copy_only=False).view(np.uint64), 8
    if pa.types.is_int64(t) or pa.types.is_uint64(t):
        return arr.to_numpy(zero_copy_only=False).view(np.uint64), 8
    raise TypeError(f"key type {t} does not pack into 64 bits")


def pack_keys(batch: pa.RecordBatch, keys: list[int]) -> np.ndarray:
    """rowkeys.cuh: pack_key -- one 4- or 8-byte column, or two 4-byte columns as hi:lo."""
    if len(keys) == 1:
        return _as_u64(batch.column(keys[0]))[0]
    if len(keys) == 2:
        (a, wa), (b, wb) = _as_u64(batch.column(keys[0])), _as_u64(batch.column(keys[1]))
        if wa == 4 and wb == 4:
            return (a << np.uint64(32)) | b
    raise TypeError("keys do not pack into 64 bits")


def partition_ids(batch: pa.RecordBatch, keys: list[int], n_parts: int) -> np.ndarray:
    """partition.cu: partition_of(fmix64(pack_key(row)), n_parts) -- the high 32 hash bits scaled to [0, n)."""
    h = fmix64(pack_keys(batch, keys))
    return (((h >> np.uint64(32)) * np.uint64(n_parts)) >> np.uint64(32)).astype(np.int64)


def round_robin(batches: list, rank: int, world: int) -> list:
    """The batches of a relation that rank `rank` of `world` scans."""
    return [b for i, b in enumerate(batches) if i % world == rank]
