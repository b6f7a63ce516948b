"""Static sharding of messages / leaves across the GPUs of one box.

Messages are independent: rank g scans messages [lo, hi) and writes its own slice of the result
words -- no collective.  Leaves shard on boundaries that are multiples of 2^block_log2 so that every
rank's block roots are level-`block_log2` nodes of the single tree; one all-gather of those 32-byte
roots and a fold of the gathered list give the global root."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int, align: int = 1):
    """Contiguous [lo, hi) of rank; every boundary except n itself is a multiple of `align`."""
    blocks = (n + align - 1) // align
    lo_b = blocks * rank // world
    hi_b = blocks * (rank + 1) // world
    return min(lo_b * align, n), min(hi_b * align, n)


def blocks_of(lo: int, hi: int, block_log2: int) -> int:
    return ((hi - lo) + (1 << block_log2) - 1) >> block_log2
