"""Counter-based synthetic activation traces for BASELINE config 5 (TEST / BENCH INFRASTRUCTURE —
only tests/, bench.py and __graft_entry__.smoke() import this package).

SURVEY.md §8d asks for C5's traces (100k test x 1.28M train x 2048-d, 1000 classes, bf16 storage) to be
generated ON THE DEVICE from a counter-based RNG, so that no 10 GB host transfer is needed, and for
the oracle to be checked on a slice regenerated on the host with the same RNG.  Everything here is
integer arithmetic in torch int64 tensors (wrapping multiply / xor / shifts: SplitMix64 of the
element's global counter), so the SAME function yields the SAME bits on `cpu` and on `cuda`:

    trace[row, k] = bf16( centre[class(row)][k] + noise(row, k) )       stored widened to float32
    class(row)    = row % classes                      (every class dealt round-robin: 1280 rows per class)
    noise         = (b0 + b1 + b2 + b3 - 510) / 128    b_i = bytes of the hash: Irwin-Hall, std ~1.15
    centre        = (b0 + b1 - 255) / 256              std ~0.41, one hash per (class, k)

Both summands are multiples of 2^-8 below 8 in magnitude, so their float32 sum is exact and the
only rounding is the final round-to-nearest-even to bf16 (identical on both devices).
"""
from __future__ import annotations

import torch

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLDEN = _s64(0x9E3779B97F4A7C15)
_MUL1 = _s64(0xBF58476D1CE4E5B9)
_MUL2 = _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """logical shift right of int64 lanes (torch's >> is arithmetic)"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(counter: torch.Tensor) -> torch.Tensor:
    z = counter + _GOLDEN
    z = (z ^ _lsr(z, 30)) * _MUL1
    z = (z ^ _lsr(z, 27)) * _MUL2
    return z ^ _lsr(z, 31)


def _bytes_sum(h: torch.Tensor, n: int) -> torch.Tensor:
    s = h & 0xFF
    for i in range(1, n):
        s = s + (_lsr(h, 8 * i) & 0xFF)
    return s


def centres(classes: int, d: int, seed: int, device) -> torch.Tensor:
    c = torch.arange(classes, dtype=torch.int64, device=device)[:, None] * d + torch.arange(d, dtype=torch.int64, device=device)
    h = splitmix64(c + _s64(seed * 0x100000001B3 + 0x51ED270B))
    return (_bytes_sum(h, 2) - 255).to(torch.float32) / 256.0


def traces(rows: torch.Tensor, d: int, classes: int, seed: int, stream: int, centre: torch.Tensor = None) -> torch.Tensor:
    """rows: int64 global row numbers (any device).  stream: 0 = training set, 1 = test set.
    Returns float32 [len(rows), d] holding bf16-representable values."""
    dev = rows.device
    if centre is None:
        centre = centres(classes, d, seed, dev)
    cnt = rows.to(torch.int64)[:, None] * d + torch.arange(d, dtype=torch.int64, device=dev)
    h = splitmix64(cnt + _s64((seed * 2 + stream + 1) * 0x9E3779B1 + (stream << 61)))
    noise = (_bytes_sum(h, 4) - 510).to(torch.float32) / 128.0
    v = centre[rows % classes] + noise
    return v.to(torch.bfloat16).to(torch.float32)


def labels(rows: torch.Tensor, classes: int) -> torch.Tensor:
    return rows % classes


def fill(out: torch.Tensor, row0: int, d: int, classes: int, seed: int, stream: int, chunk: int = 16384):
    """out[i] = trace of global row row0 + i, generated in chunks (bounded temporaries)."""
    dev = out.device
    centre = centres(classes, d, seed, dev)
    for s in range(0, out.shape[0], chunk):
        e = min(out.shape[0], s + chunk)
        rows = torch.arange(row0 + s, row0 + e, dtype=torch.int64, device=dev)
        out[s:e] = traces(rows, d, classes, seed, stream, centre)
    return out
