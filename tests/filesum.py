"""Position-aware additive checksum of a file of 64-bit words (test infrastructure).

    s0 = sum_j w_j                                   (mod 2^64)
    s1 = sum_j w_j * (2 j + 1)                       (mod 2^64)
    s2 = sum_j (w_j ^ (w_j >> 23)) * (G * (j + 1) | 1)   (mod 2^64),  G = 0x9E3779B97F4A7C15

`j` is the GLOBAL index of the word in the file, so the checksum of a file is the sum of the checksums of
any tiling of it into ranges: every rank of a sharded build computes the sums of its own [node_lo, node_hi)
records (on the device, `filesum_torch`) and the totals are compared with the value the golden generator
computed from the reference's file (`filesum_file`). A moved, missing or altered word changes s1 / s2.
"""
from __future__ import annotations

import numpy as np

G = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


def filesum_array(words: np.ndarray, first: int = 0) -> tuple[int, int, int]:
    """words: uint64 array (any shape, C order); first: global index of words[0]."""
    w = np.ascontiguousarray(words).reshape(-1).view(np.uint64)
    s0 = s1 = s2 = 0
    step = 1 << 24
    with np.errstate(over="ignore"):
        for lo in range(0, w.size, step):
            c = w[lo:lo + step]
            j = np.arange(first + lo, first + lo + c.size, dtype=np.uint64)
            s0 += int(c.sum(dtype=np.uint64))
            s1 += int((c * (j * np.uint64(2) + np.uint64(1))).sum(dtype=np.uint64))
            s2 += int(((c ^ (c >> np.uint64(23))) * (((j + np.uint64(1)) * np.uint64(G)) | np.uint64(1))).sum(dtype=np.uint64))
    return s0 & MASK, s1 & MASK, s2 & MASK


def filesum_file(path: str) -> tuple[int, int, int]:
    s = [0, 0, 0]
    first = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(256 << 20)
            if not b:
                break
            assert len(b) % 8 == 0
            a = np.frombuffer(b, dtype=np.uint64)
            p = filesum_array(a, first)
            s = [(x + y) & MASK for x, y in zip(s, p)]
            first += a.size
    return tuple(s)


def filesum_torch(words, first: int = 0) -> tuple[int, int, int]:
    """words: torch int64 tensor on any device (the bit pattern of the uint64 words); two's-complement int64
    arithmetic wraps exactly like uint64 arithmetic."""
    import torch
    w = words.reshape(-1)
    s0 = s1 = s2 = 0
    step = 1 << 26
    g = G - (1 << 64)                                     # the same bit pattern as a signed value
    for lo in range(0, w.numel(), step):
        c = w[lo:lo + step]
        j = torch.arange(first + lo, first + lo + c.numel(), dtype=torch.int64, device=w.device)
        s0 += int(c.sum())
        s1 += int((c * (j * 2 + 1)).sum())
        sh = (c >> 23) & ((1 << 41) - 1)                  # logical shift
        s2 += int(((c ^ sh) * (((j + 1) * g) | 1)).sum())
    return s0 & MASK, s1 & MASK, s2 & MASK


def add(a, b):
    return tuple((x + y) & MASK for x, y in zip(a, b))
