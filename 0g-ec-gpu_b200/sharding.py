"""Contiguous sharding of one MSM over several devices / ranks: the partition of
MultiexpKernel::parallel_multiexp (ec-gpu-proxy/src/multiexp.rs:329-337)."""
from __future__ import annotations


def chunk_size(n: int, parts: int) -> int:
    """ceil(n / parts): the maximum number of terms per device (multiexp.rs:333-334)."""
    if parts <= 0:
        raise ValueError("parts must be positive")
    return (n + parts - 1) // parts


def shard_range(n: int, parts: int, index: int) -> tuple[int, int]:
    """[start, end) of shard `index`; trailing shards may be empty, as with `chunks()` in Rust."""
    c = chunk_size(n, parts)
    start = min(index * c, n)
    return start, min(start + c, n)
