"""Device-side orchestration of the scoring kernels (PyTorch = memory, streams, NCCL plumbing).

Two engines sit under the reference-compatible classes in `simple_tip_b200.core`:

* `NnEngine`  — class-grouped training traces resident in HBM (original dtype for the exact
  re-rank + packed bf16 operand for the tensor-core filter); `search()` runs
  filter -> re-rank for a batch of class-sorted queries.  Used twice per DSA call
  (surprise.py:615-631: same-class nearest neighbour, then other-class nearest neighbour
  of the winning TRAIN row).
* `KdeEngine` — whitened training traces as a 3-segment split-bf16 operand; `log_kernel_sum()`
  returns per-query (max, sum) of exp(-|p_i - q_j|^2 / 2) in the log domain
  (scipy 1.4.1 gaussian_kernel_estimate, called from stable_kde.py:101).

The pure-host planning helpers (`class_layout`, `build_items`, `shard_rows`) and the
cross-rank reduction protocol (`TrainShardComm`) contain no CUDA calls and are unit-tested on
CPU (gloo, world_size 2).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

POOL_FRAC = float(os.environ.get("B200TIP_POOL_FRAC", "0.2"))   # share of the filter's work handed out dynamically
SPECULATIVE = os.environ.get("B200TIP_SPECULATIVE", "1") != "0"   # graph plans without exhaustive-scan launches
POOL_MIN_TILES = 32       # tiles per CTA below which work lists stay purely static
POOL_TILES = int(os.environ.get("B200TIP_POOL_TILES", "4"))      # train tiles per dynamic item
DEFAULT_CAP = 64          # candidate chunks per query (short traces)
DEFAULT_CAP_LONG = 256    # long traces: distances concentrate, more rows fall inside the window
SEEDS = os.environ.get("B200TIP_SEEDS", "1") != "0"      # fit-time upper bounds that seed stage 2's running minima

# bench.py sets this to a list to collect (kernel name, algorithmic flops, start event, end event)
# around the tensor-core launches; None (default) records nothing.
PROFILE = None


# ------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------
def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("simple_tip_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def tip_dtype(dt) -> int:
    if dt in (torch.float32, np.float32, np.dtype("float32")):
        return _lib.TIP_F32
    if dt in (torch.float64, np.float64, np.dtype("float64")):
        return _lib.TIP_F64
    raise TypeError(f"activation traces must be float32 or float64, got {dt}")


def to_device(a, dev: torch.device) -> torch.Tensor:
    """Host -> HBM.  Pinned sources (e.g. views of torch pinned tensors) copy asynchronously.
    Traces that already live on the GPU (torch CUDA tensors) are used in place."""
    if isinstance(a, torch.Tensor):
        return a.to(dev, non_blocking=True).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)


def device_matrix(layers) -> Optional[torch.Tensor]:
    """Activation traces handed over as device tensors (one [N, ...] tensor or a list of them, the
    form a forward hook produces): the flattened [N, D] matrix on the GPU, else None.  Mirrors
    `_flatten_layers` (surprise.py:62-66) without leaving HBM."""
    if isinstance(layers, torch.Tensor):
        return layers.reshape(layers.shape[0], -1) if layers.is_cuda else None
    if isinstance(layers, (list, tuple)) and len(layers) > 0 and all(isinstance(l, torch.Tensor) and l.is_cuda
                                                                     for l in layers):
        flat = [l.reshape(l.shape[0], -1) for l in layers]
        return flat[0] if len(flat) == 1 else torch.cat(flat, dim=1)
    return None


def host_array(x) -> np.ndarray:
    """Labels / predictions as a NumPy array whatever they came as (device tensors are copied)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return x


NP_DTYPE = {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
            torch.float16: np.dtype(np.float16), torch.bfloat16: None}


# ------------------------------------------------------------------------------------------
# host-side planning (no CUDA)
# ------------------------------------------------------------------------------------------
def class_layout(labels: np.ndarray, num_classes: int) -> Tuple[np.ndarray, np.ndarray]:
    """Stable grouping by class.  Returns (order, offsets[num_classes+1]); `order` lists the
    positions of rows with 0 <= label < num_classes, grouped by class, ascending position inside
    a class (the order `np.argwhere(pred == label)` gives, surprise.py:554,581)."""
    labels = np.asarray(labels)
    ok = (labels >= 0) & (labels < num_classes)
    # NumPy's stable sort is a radix sort for <= 16-bit integers: ~10x faster than on int64
    small = np.uint16 if num_classes <= 65535 else np.int64
    if ok.all():
        keys = labels.astype(small)
        order = np.argsort(keys, kind="stable")
        counts = np.bincount(keys, minlength=num_classes)[:num_classes]
    else:
        valid = np.flatnonzero(ok)
        keys = labels[valid].astype(small)
        order = valid[np.argsort(keys, kind="stable")]
        counts = np.bincount(keys, minlength=num_classes)[:num_classes]
    offsets = np.zeros(num_classes + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return order, offsets


def shard_rows(labels: np.ndarray, num_classes: int, rank: int, world: int) -> np.ndarray:
    """Rows of the training set kept by `rank`: every class is dealt round-robin, so each rank
    holds ~N_c/world rows of every class in ascending original order."""
    order, off = class_layout(labels, num_classes)
    keep = []
    for c in range(num_classes):
        rows = order[off[c]:off[c + 1]]
        keep.append(rows[rank::world])
    return np.sort(np.concatenate(keep)) if keep else np.zeros(0, dtype=np.int64)


def span_tiles_for(total_tile_pairs: int, sms: int, per_sm: int = 6, lo: int = 2, hi: int = 16) -> int:
    """Column tiles per work item: aim for ~per_sm items per SM so the static round-robin balances."""
    return int(min(hi, max(lo, round(total_tile_pairs / max(1, sms * per_sm)))))


def build_items(q_off: np.ndarray, ranges_per_class: Sequence[Sequence[Tuple[int, int]]], span_tiles: int,
                row_tile: int = _lib.ROW_TILE, col_tile: int = _lib.COL_TILE) -> Tuple[np.ndarray, int]:
    """Work list of the tensor-core pass.  q_off[c]..q_off[c+1] are the (class-sorted) query rows
    of class c; ranges_per_class[c] the train-row ranges they must be compared with.  Every item is
    one row_tile-row query tile x one span of at most span_tiles*col_tile train rows.  Returns
    (int32 array [n, 6] laid out as tip_work_item, max number of spans per class)."""
    width = span_tiles * col_tile
    out = []
    max_slots = 1
    for c, ranges in enumerate(ranges_per_class):
        cnt = int(q_off[c + 1] - q_off[c])
        if cnt <= 0:
            continue
        spans = []
        for lo, hi in ranges:
            for s in range(int(lo), int(hi), width):
                spans.append((s, min(s + width, int(hi))))
        if not spans:
            continue
        max_slots = max(max_slots, len(spans))
        r0 = np.arange(int(q_off[c]), int(q_off[c + 1]), row_tile, dtype=np.int64)
        rows = np.minimum(row_tile, int(q_off[c + 1]) - r0)
        sp = np.asarray(spans, dtype=np.int64)
        nt, ns = r0.shape[0], sp.shape[0]
        item = np.zeros((nt * ns, 6), dtype=np.int32)
        item[:, 0] = np.repeat(r0, ns)
        item[:, 1] = np.repeat(rows, ns)
        item[:, 2] = np.tile(sp[:, 0], nt)
        item[:, 3] = np.tile(sp[:, 1], nt)
        item[:, 4] = np.tile(np.arange(ns), nt)
        out.append(item)
    if not out:
        return np.zeros((0, 6), dtype=np.int32), max_slots
    return np.concatenate(out), max_slots


def build_other_class_items(q_off: np.ndarray, t_off: np.ndarray, row_tile: int, col_tile: int, span_of) -> np.ndarray:
    """Work list of DSA's second stage (every query against the train rows of all OTHER classes,
    surprise.py:622-629).  Queries are tiled in class-sorted order regardless of class boundaries, so
    tiles are full however many classes there are.  A tile whose queries share one class scans the
    two train ranges around that class; a tile that mixes classes scans every train row and each
    query masks out its own class in the epilogue (item flag bit 0)."""
    q_off = np.asarray(q_off, dtype=np.int64)
    t_off = np.asarray(t_off, dtype=np.int64)
    m, n = int(q_off[-1]), int(t_off[-1])
    if m == 0 or n == 0:
        return np.zeros((0, 6), dtype=np.int32)
    starts = np.arange(0, m, row_tile, dtype=np.int64)
    rows = np.minimum(row_tile, m - starts)
    first_cls = np.searchsorted(q_off, starts, side="right") - 1
    last_cls = np.searchsorted(q_off, starts + rows - 1, side="right") - 1
    single = first_cls == last_cls
    cols = np.where(single, n - (t_off[np.minimum(first_cls + 1, len(t_off) - 1)] - t_off[first_cls]), n)
    pairs = int(np.sum(np.ceil(cols / col_tile)))
    width = span_of(pairs) * col_tile
    out = []
    for r0, nr, c, one in zip(starts, rows, first_cls, single):
        ranges = [(0, int(t_off[c])), (int(t_off[c + 1]), n)] if one else [(0, n)]
        spans = [(s0, min(s0 + width, hi)) for lo, hi in ranges for s0 in range(lo, hi, width)]
        if not spans:
            continue
        item = np.zeros((len(spans), 6), dtype=np.int32)
        item[:, 0], item[:, 1] = r0, nr
        item[:, 2:4] = np.asarray(spans, dtype=np.int64)
        item[:, 4] = np.arange(len(spans))
        item[:, 5] = 0 if one else 1
        out.append(item)
    return np.concatenate(out) if out else np.zeros((0, 6), dtype=np.int32)


def query_tiles(q_off: np.ndarray, ranges_per_class, t_off: Optional[np.ndarray], row_tile: int, mixed: bool):
    """Query tiles of one search stage as (first row, rows, train ranges, flag) tuples.
    mixed=False: class-aligned tiles (a tile never crosses a class boundary) scanning the ranges of
    their class.  mixed=True (DSA's other-class stage with few queries per class): tiles cut
    regardless of class boundaries; a tile that mixes classes scans every train row and each query
    masks its own class in the epilogue (flag bit 0)."""
    q_off = np.asarray(q_off, dtype=np.int64)
    tiles = []
    if not mixed:
        for c, ranges in enumerate(ranges_per_class):
            rs = [(int(lo), int(hi)) for lo, hi in ranges if int(hi) > int(lo)]
            if not rs:
                continue
            for r0 in range(int(q_off[c]), int(q_off[c + 1]), row_tile):
                tiles.append((r0, min(row_tile, int(q_off[c + 1]) - r0), rs, 0))
        return tiles
    t_off = np.asarray(t_off, dtype=np.int64)
    m, n = int(q_off[-1]), int(t_off[-1])
    for r0 in range(0, m, row_tile):
        nr = min(row_tile, m - r0)
        c0 = int(np.searchsorted(q_off, r0, side="right") - 1)
        c1 = int(np.searchsorted(q_off, r0 + nr - 1, side="right") - 1)
        if c0 == c1:
            rs = [(lo, hi) for lo, hi in ((0, int(t_off[c0])), (int(t_off[c0 + 1]), n)) if hi > lo]
            flag = 0
        else:
            rs, flag = ([(0, n)] if n > 0 else []), 1
        if rs:
            tiles.append((r0, nr, rs, flag))
    return tiles


def build_balanced_items(tiles, col_tile: int, n_cta: int, item_cost: float = 0.75, pool_frac: float = 0.0,
                         pool_tiles: int = 4) -> Tuple[np.ndarray, int]:
    """Work list for the persistent resident-query kernel.  Returns (items [n, 6], n_static).

    Static part, balanced for the kernel's schedule (CTA b runs items b, b + G, b + 2G, ...;
    G = min(n_cta, #items)): the (query tile x train tile) pairs are linearised query-tile-major
    and cut into G contiguous chunks of equal cost, so every CTA streams the same number of train
    tiles (+-1) and loads as few query tiles as possible; a chunk becomes one item per (query tile,
    train range) it touches.  `item_cost` is the price of starting an item (query-tile load +
    pipeline refill) in train-tile units.

    Dynamic part: the last `pool_frac` of every chunk is cut off and split into items of at most
    `pool_tiles` train tiles that follow the static items in the array; CTAs pull them through an
    atomic counter when their static share is done, which absorbs run-time differences between
    CTAs (data-dependent candidate handling) that no static cut can foresee."""
    segs = []          # (q_row0, q_rows, lo, hi, flag, ntiles)
    for r0, nr, ranges, flag in tiles:
        for lo, hi in ranges:
            nt = -(-(int(hi) - int(lo)) // col_tile)
            if nt > 0:
                segs.append((int(r0), int(nr), int(lo), int(hi), int(flag), nt))
    if not segs:
        return np.zeros((0, 6), dtype=np.int32), 0
    total = sum(sg[5] for sg in segs)
    g = int(min(n_cta, total))
    # Cost axis: every segment costs item_cost up front, then 1 per train tile.  A tile belongs to
    # the CTA whose equal share of the axis holds the tile's centre; consecutive tiles of a
    # segment with the same owner form one piece.
    cost_total = total + item_cost * len(segs)
    per_cta = [[] for _ in range(g)]           # pieces: (segment index, first tile, one past last tile)
    pos = 0.0
    for si, (r0, nr, lo, hi, flag, nt) in enumerate(segs):
        pos += item_cost
        owner = np.minimum(((pos + np.arange(nt) + 0.5) * (g / cost_total)).astype(np.int64), g - 1)
        cuts = np.flatnonzero(np.diff(owner)) + 1
        starts = np.concatenate(([0], cuts))
        ends = np.concatenate((cuts, [nt]))
        for t0, t1 in zip(starts, ends):
            per_cta[int(owner[t0])].append((si, int(t0), int(t1)))
        pos += nt

    def as_item(si, t0, t1):
        r0, nr, lo, hi, flag, _ = segs[si]
        return (r0, nr, lo + t0 * col_tile, min(hi, lo + t1 * col_tile), 0, flag)

    pool, tails = [], []
    if pool_frac > 0.0 and total >= POOL_MIN_TILES * g:   # short launches: the item-start cost outweighs the tail
        for lst in per_cta:
            give = int(round(pool_frac * sum(t1 - t0 for _, t0, t1 in lst)))
            tail = []
            while give > 0 and lst:
                si, t0, t1 = lst[-1]
                take = min(give, t1 - t0)
                tail.append((si, t1 - take, t1))
                give -= take
                if take == t1 - t0:
                    lst.pop()
                else:
                    lst[-1] = (si, t0, t1 - take)
            mine = []
            for si, t0, t1 in reversed(tail):             # keep the train order inside a chunk
                for p0 in range(t0, t1, pool_tiles):
                    mine.append(as_item(si, p0, min(t1, p0 + pool_tiles)))
            tails.append(mine)
        # Pool order: round-robin over the chunks.  CTAs arrive at the pool at about the same time and
        # take one item per round, so a CTA's successive draws tend to be successive pieces of one
        # chunk's tail — same query tile, which the kernel then keeps resident.
        for r in range(max((len(tl) for tl in tails), default=0)):
            pool.extend(tl[r] for tl in tails if len(tl) > r)
    per_cta = [[as_item(*pc) for pc in lst] for lst in per_cta if lst]
    per_cta.sort(key=len, reverse=True)       # CTAs with more items first: rounds stay dense prefixes
    g = len(per_cta)
    rounds = len(per_cta[0]) if per_cta else 0
    out = []
    empty = (0, 0, 0, 0, 0, 0)                # col0 == col1: the kernel skips it
    for r in range(rounds):
        row = [lst[r] for lst in per_cta if len(lst) > r]
        if r + 1 < rounds or pool:
            row += [empty] * (g - len(row))   # keep item index = round * G + CTA
        out.extend(row)
    n_static = len(out)
    out.extend(pool)
    return np.asarray(out, dtype=np.int32).reshape(-1, 6), n_static


def span_major(items: np.ndarray) -> np.ndarray:
    """Order the streaming kernel's work list by (train span, query tile).  CTA b runs items b, b + G, ...,
    so the CTAs that run at the same time then scan the SAME train rows for different query tiles: a train
    tile comes out of HBM once and is shared through L2.  Query-tile-major order re-reads the whole packed
    training set from HBM for every 128-row query tile — at C5's size (2.7 GB per rank) that made the
    tensor-core pass HBM-bound at 128 flop/B (measured: 42 ms instead of 26 for 10k x 640k x 2048)."""
    if items.shape[0] < 2:
        return items
    return np.ascontiguousarray(items[np.lexsort((items[:, 0], items[:, 2]))])


def count_tile_pairs(q_off: np.ndarray, ranges_per_class, row_tile: int = _lib.ROW_TILE,
                     col_tile: int = _lib.COL_TILE) -> int:
    total = 0
    for c, ranges in enumerate(ranges_per_class):
        cnt = int(q_off[c + 1] - q_off[c])
        if cnt <= 0:
            continue
        cols = sum(max(0, int(hi) - int(lo)) for lo, hi in ranges)
        total += math.ceil(cnt / row_tile) * math.ceil(cols / col_tile)
    return total


# ------------------------------------------------------------------------------------------
# cross-rank protocol (N_train sharded over the ranks of one NVSwitch box)
# ------------------------------------------------------------------------------------------
KEY_NONE_HI32 = 0x7F800000           # +inf as float bits: "this shard has no row of that range"
KEY_NONE_LO = 0x7FFFFFFF
P2P_MIN_RECORDS = 1 << 18            # receive slots per rank: 2 x world x 4 MiB


def pack_winner_keys(dist: torch.Tensor, gid: torch.Tensor) -> torch.Tensor:
    """float32 distances (>= 0, NaN = no row on this shard) and original indices -> int64 keys
    (float bits << 32 | index).  Non-negative IEEE floats order like their bit patterns and the
    sign bit is clear, so ONE integer MIN all-reduce yields np.min and np.argmin's first occurrence
    (surprise.py:645-647) over all shards.  Plain torch ops: runs on CPU (gloo test) and CUDA."""
    assert dist.dtype == torch.float32
    none = torch.isnan(dist) | (gid < 0)
    hi = torch.where(none, torch.full_like(dist, float("inf")), dist).view(torch.int32).to(torch.int64)
    lo = torch.where(none, torch.full_like(gid, KEY_NONE_LO, dtype=torch.int64), gid.to(torch.int64))
    return (hi << 32) | lo


def unpack_winner_keys(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of pack_winner_keys: (float32 distance with NaN for 'no row anywhere', int64 index or -1)."""
    hi = (keys >> 32).to(torch.int32)
    lo = keys & 0xFFFFFFFF
    none = (hi >= KEY_NONE_HI32) | (lo >= KEY_NONE_LO)
    dist = torch.where(none, torch.full((), float("nan"), dtype=torch.float32, device=keys.device), hi.view(torch.float32))
    return dist, torch.where(none, torch.full_like(lo, -1), lo)


class P2PExchange:
    """Symmetric receive buffers of libb200tip's tip_comm, one per rank, opened across the
    processes of one box with CUDA IPC; the handles travel through torch.distributed once."""

    def __init__(self, dist, group, dev: torch.device, cap_records: int):
        self.lib = _lib.load()
        self.cap = int(cap_records)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.handle = None
        self.local = C.c_void_p(0)
        handle = (C.c_ubyte * 64)()
        rc = self.lib.tip_comm_alloc(self.world, self.cap, C.byref(self.local), handle)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        err = None
        if int(ok.item()) == 1:
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
            out = C.c_void_p(0)
            rc = self.lib.tip_comm_open(self.rank, self.world, self.local, blob, self.cap, C.byref(out))
            if rc == 0:
                self.handle = out
            else:
                err = self.lib.tip_last_error().decode(errors="replace")
        else:
            err = "tip_comm_alloc failed on some rank"
        ok = torch.tensor([1 if self.handle is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # all ranks take the same path
        if int(ok.item()) != 1:
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable: {err or 'a peer could not map the buffers'}")

    def close(self):
        if self.handle is not None:
            self.lib.tip_comm_close(self.handle)
            self.handle = None
        if self.local:
            self.lib.tip_comm_free_local(self.local)
            self.local = C.c_void_p(0)

    def push_nn(self, dist_t: torch.Tensor, gid: Optional[torch.Tensor]):
        _lib.check(self.lib.tip_comm_push_nn(self.handle, _p(dist_t), tip_dtype(dist_t.dtype), _p(gid), dist_t.shape[0],
                                             _stream()), "tip_comm_push_nn")

    def min_into(self, out: torch.Tensor):
        _lib.check(self.lib.tip_comm_min(self.handle, tip_dtype(out.dtype), out.shape[0], _p(out), _stream()), "tip_comm_min")

    def lse(self, mx: torch.Tensor, sm: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        m = mx.shape[0]
        _lib.check(self.lib.tip_comm_push_lse(self.handle, _p(mx), _p(sm), m, _stream()), "tip_comm_push_lse")
        gm = torch.empty(m, dtype=torch.float32, device=mx.device)
        gs = torch.empty(m, dtype=torch.float64, device=mx.device)
        _lib.check(self.lib.tip_comm_lse(self.handle, m, _p(gm), _p(gs), _stream()), "tip_comm_lse")
        return gm, gs


class TrainShardComm:
    """Exchange steps of the N_train-sharded search over a torch.distributed group.

    On GPUs of one NVSwitch box the per-query records travel as peer stores fused into our own
    kernels (`P2PExchange`, csrc/shard.cu); otherwise — gloo in the CPU tests, B200TIP_EXCHANGE=nccl,
    or no peer access — as ONE integer MIN all-reduce per stage through torch.distributed.  All
    messages are O(N_test) scalars; the stage-2 query rows are gathered from the replicated raw
    training set by original index (or, without a replica, summed from their owners)."""

    def __init__(self, group=None, exchange: Optional[str] = None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.exchange = exchange or os.environ.get("B200TIP_EXCHANGE", "auto")     # auto | p2p | nccl
        assert self.exchange in ("auto", "p2p", "nccl"), self.exchange
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.collectives = 0          # torch.distributed all-reduces issued on the data path
        self._p2p = None
        self._p2p_failed = None

    # -- peer-memory exchange -----------------------------------------------------------------
    def p2p(self, dev: torch.device, records: int) -> Optional[P2PExchange]:
        """The peer-memory exchange sized for `records` queries, or None (collective fallback).
        Collective: every rank must call it with the same arguments."""
        if self.exchange == "nccl" or self.world < 2 or dev.type != "cuda" or self.world > 16:
            return None
        if self._p2p is not None:
            # never re-allocated: captured graphs hold its addresses.  Larger batches take the collective path.
            return self._p2p if self._p2p.cap >= records else None
        if self._p2p_failed is not None:
            return None
        try:
            self._p2p = P2PExchange(self.dist, self.group, dev, max(int(records), P2P_MIN_RECORDS))
        except (RuntimeError, _lib.TipError) as e:
            if self.exchange == "p2p":
                raise
            self._p2p_failed = str(e)
            return None
        return self._p2p

    def close(self):
        if self._p2p is not None:
            self._p2p.close()
            self._p2p = None

    def min_(self, t: torch.Tensor) -> torch.Tensor:
        self.collectives += 1
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        self.collectives += 1
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def sum_(self, t: torch.Tensor) -> torch.Tensor:
        self.collectives += 1
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    # -- DSA stage 1: global (distance, lowest original index) and the winning rows -----------
    def reduce_winner_index(self, dist_a: torch.Tensor, gid: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global minimum distance (NaN where no shard has a row of the class) and the original
        index of the winner — lowest index among exact ties, i.e. np.argmin's first occurrence.
        float32: one MIN all-reduce of packed 64-bit keys; float64: MIN of the distances, then MIN
        of the index among the ranks that hold the minimum."""
        if dist_a.dtype == torch.float32:
            return unpack_winner_keys(self.min_(pack_winner_keys(dist_a, gid)))
        big = torch.iinfo(torch.int64).max
        local = torch.where(torch.isnan(dist_a), torch.full_like(dist_a, float("inf")), dist_a)
        gmin = self.min_(local.clone())
        cand = torch.where((local == gmin) & (gid >= 0), gid.to(torch.int64),
                           torch.full_like(gid, big, dtype=torch.int64))
        ggid = self.min_(cand)
        none = torch.isinf(gmin) | (ggid == big)
        return (torch.where(none, torch.full_like(gmin, float("nan")), gmin),
                torch.where(none, torch.full_like(ggid, -1), ggid))

    def reduce_winners(self, dist_a: torch.Tensor, gid: torch.Tensor, rows: Optional[torch.Tensor] = None,
                       train_full: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """dist_a[m] local minima (NaN = this shard has no row of the class), gid[m] original
        index of the local winner.  Returns the global minimum distance, the global winner's original
        index and its trace on every rank: gathered from `train_full` (the replicated training set in
        original order) when given, else all-reduced from `rows[m, d]` (owner contributes, others
        zeros -> exact)."""
        gmin, ggid = self.reduce_winner_index(dist_a, gid)
        if train_full is not None:
            winners = train_full.index_select(0, ggid.clamp(min=0))
            winners = torch.where((ggid >= 0)[:, None], winners, torch.zeros_like(winners))
        else:
            mine = (ggid >= 0) & (gid.to(torch.int64) == ggid)
            winners = torch.where(mine[:, None], rows, torch.zeros_like(rows))
            self.sum_(winners)
        return gmin, ggid, winners

    def reduce_min_nan(self, d: torch.Tensor) -> torch.Tensor:
        """min over ranks where NaN means 'empty range on this shard'."""
        local = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d)
        g = self.min_(local)
        return torch.where(torch.isinf(g), torch.full_like(g, float("nan")), g)

    # -- LSA: merge per-shard (max, sum) -------------------------------------------------------
    def reduce_lse(self, mx: torch.Tensor, sm: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        g = self.max_(mx.clone())
        scale = torch.where(torch.isinf(mx) & (mx < 0), torch.zeros_like(mx), torch.exp(mx - g))
        s = self.sum_((sm * scale).to(torch.float64))
        return g, s


# ------------------------------------------------------------------------------------------
# nearest-neighbour engine (DSA)
# ------------------------------------------------------------------------------------------
class NnEngine:
    def __init__(self, t_sorted: torch.Tensor, class_off: np.ndarray, t_gid: torch.Tensor, cap: Optional[int] = None,
                 seeds: bool = True):
        """t_sorted: [n, d] float32/float64 on the GPU, rows grouped by class (offsets class_off,
        ascending original index inside a class); t_gid[n] int32 original indices."""
        self.dev = require_cuda()
        self.lib = _lib.load()
        self.t = t_sorted.contiguous()
        self.n, self.d = self.t.shape
        self.dtype = tip_dtype(self.t.dtype)
        self.class_off = np.asarray(class_off, dtype=np.int64)
        self.num_classes = self.class_off.shape[0] - 1
        self.class_off_dev = torch.from_numpy(self.class_off.astype(np.int32)).to(self.dev)
        self.t_gid = t_gid.to(torch.int32).contiguous()
        self.cap = int(cap) if cap else (DEFAULT_CAP if self.d <= 256 else DEFAULT_CAP_LONG)
        self.pitch = int(self.lib.tip_pair_pitch(self.d, 1))
        sms = C.c_int(0)
        _lib.check(self.lib.tip_device_info(C.byref(sms), None, None), "tip_device_info")
        self.sms = sms.value
        qt, tt = C.c_int32(0), C.c_int32(0)
        _lib.check(self.lib.tip_nn_filter_tile(self.d, C.byref(qt), C.byref(tt)), "tip_nn_filter_tile")
        self.row_tile, self.col_tile = qt.value, tt.value
        self.resident = int(self.lib.tip_nn_filter_kind(self.d)) == 1     # resident-query kernel (short traces)
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self._item_cache = {}
        self._plans = {}
        self._work = {}
        self._inputs = {}
        self._capture_refs = None     # while a plan is being captured: engine-owned tensors its graph reads
        self.sched_counter = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.last_cand_cnt_by_mode = {}
        # N_train-sharded engines: the whole (post-subsample) training set in ORIGINAL row order, replicated on
        # every rank, from which the global stage-1 winners are gathered by index (set by the owner)
        self.t_full = None
        if self.n > 0:
            # Distances are translation invariant: the packed operands hold bf16(x - center).  For general traces the
            # global mean shrinks the rounding error; traces that are exactly representable in bf16 (bf16-stored
            # activations, BASELINE config 5) are packed LOSSLESSLY with center = 0 — centring would only create a
            # rounding error (at C5: 33 candidate chunks per query instead of ~3, and a 7 ms re-rank per pass).
            if self._bf16_exact(self.t):
                self.center = torch.zeros(self.d, dtype=torch.float32, device=self.dev)
            else:
                self.center = self.t.mean(dim=0, dtype=torch.float64).to(torch.float32).contiguous()
            self.t_pack = torch.empty((self.n, self.pitch), dtype=torch.bfloat16, device=self.dev)
            sq = torch.empty(self.n, dtype=torch.float32, device=self.dev)
            err = torch.empty(self.n, dtype=torch.float32, device=self.dev)
            _lib.check(self.lib.tip_pair_prep(_p(self.t), self.dtype, self.n, self.d, _p(self.center), _lib.ROLE_TRAIN,
                                              1, -2.0, 1.0, _p(self.t_pack), _p(sq), _p(err), _stream()), "tip_pair_prep")
            self.rmax = float(torch.sqrt(sq.max()).item()) * (1.0 + 1e-6)
            self.errmax = float(err.max().item()) * (1.0 + 1e-6)   # measured bf16 rounding of the train rows
        else:
            self.center = torch.zeros(self.d, dtype=torch.float32, device=self.dev)
            self.t_pack, self.rmax, self.errmax = None, 0.0, 0.0
        # Stage-2 seeds (short traces / resident-query kernel, where the warm-up of the running minima costs ~15 % of
        # the filter): for every train row an upper bound on the distance to its nearest other-class row = the exact
        # distance to the nearest one inside a fixed stratified sample of the training set, found once at fit time
        # with this very engine machinery.
        self.seed_b = None
        if seeds and SEEDS and self.n >= 4096 and self.num_classes >= 2 and self.resident:
            self.seed_b = self._other_class_seeds()
        # Optional fit-time table (build_other_class_table): dist_b of surprise.py:622-631 depends on the test input
        # only through WHICH train row won stage 1, so it can be tabulated per train row once per training set.
        self.table_b = None

    def build_other_class_table(self, chunk_rows: int = 1 << 16) -> torch.Tensor:
        """For every train row the exact NumPy-order distance to its nearest row of another class (NaN when there is
        none): the second stage of DSA for any future winner, found with the same filter + re-rank as a scoring call.
        N_train x N_train pairs once per training set; a call then only runs stage 1 and looks dist_b up."""
        cls_of_row = np.repeat(np.arange(self.num_classes), np.diff(self.class_off)).astype(np.int32)
        table = torch.empty(self.n, dtype=self.t.dtype, device=self.dev)
        for s in range(0, self.n, chunk_rows):
            e = min(self.n, s + chunk_rows)
            q_off = (np.clip(self.class_off, s, e) - s).astype(np.int64)
            q_class = torch.from_numpy(cls_of_row[s:e]).to(self.dev)
            table[s:e] = self.search(self.t[s:e], q_class, q_off, _lib.RANGE_OTHER_CLASSES)[0]
        torch.cuda.current_stream().synchronize()
        self.table_b = table
        return table

    def _other_class_seeds(self, sample_rows: int = 8192) -> Optional[torch.Tensor]:
        step = max(1, self.n // sample_rows)
        pick = torch.arange(0, self.n, step, device=self.dev)
        cls_of_row = np.repeat(np.arange(self.num_classes), np.diff(self.class_off))
        sub_cls = cls_of_row[::step][:pick.shape[0]]
        sub_off = np.concatenate([[0], np.cumsum(np.bincount(sub_cls, minlength=self.num_classes))]).astype(np.int64)
        if np.count_nonzero(np.diff(sub_off)) < 2:
            return None
        sub = NnEngine(self.t.index_select(0, pick), sub_off, self.t_gid.index_select(0, pick), seeds=False)
        q_class = torch.from_numpy(cls_of_row.astype(np.int32)).to(self.dev)
        ub = sub.search(self.t, q_class, self.class_off, _lib.RANGE_OTHER_CLASSES)[0].to(torch.float32)
        ub = torch.where(torch.isnan(ub), torch.full_like(ub, float("inf")), ub).contiguous()
        torch.cuda.current_stream().synchronize()
        return ub

    @staticmethod
    def _bf16_exact(t: torch.Tensor, chunk_rows: int = 1 << 18) -> bool:
        """True if every trace value survives a round trip through bfloat16 (checked in row chunks)."""
        if t.dtype != torch.float32:
            return False
        for r0 in range(0, t.shape[0], chunk_rows):
            c = t[r0:r0 + chunk_rows]
            if not bool((c.to(torch.bfloat16).to(torch.float32) == c).all()):
                return False
        return True

    @classmethod
    def from_host(cls, train: np.ndarray, labels: np.ndarray, num_classes: int, gids: Optional[np.ndarray] = None,
                  cap: Optional[int] = None, seeds: bool = True) -> "NnEngine":
        dev = require_cuda()
        order, off = class_layout(labels, num_classes)
        t = to_device(train, dev)
        idx = torch.from_numpy(order).to(dev)
        gid = idx if gids is None else torch.from_numpy(np.asarray(gids)[order]).to(dev)
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float32)          # device-resident bf16 / fp16 traces: exact widening
        return cls(t.index_select(0, idx), off, gid, cap, seeds)

    def work_buffer(self, m: int, dtype: torch.dtype) -> torch.Tensor:
        """Scratch of tip_nn_rerank for m queries: zero-filled once, every call leaves it re-armed."""
        key = (int(m), dtype)
        w = self._work.get(key)
        if w is None:
            if len(self._work) >= 16:
                self._work.clear()
            nbytes = int(self.lib.tip_nn_rerank_work_bytes(m, tip_dtype(dtype)))
            w = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
            self._work[key] = w
        return w

    def input_buffer(self, rows: int, dtype: torch.dtype) -> torch.Tensor:
        """Persistent HBM landing buffer for a batch of `rows` test traces (the H2D target of
        DSA.__call__ and the gather source of the captured search)."""
        key = (int(rows), dtype)
        b = self._inputs.get(key)
        if b is None:
            if len(self._inputs) >= 8:
                self._inputs.pop(next(iter(self._inputs)))
            b = torch.zeros((rows, self.d), dtype=dtype, device=self.dev)
            self._inputs[key] = b
        return b

    def ranges(self, mode: int):
        off = self.class_off
        if mode == _lib.RANGE_SAME_CLASS:
            return [[(off[c], off[c + 1])] for c in range(self.num_classes)]
        return [[(0, off[c]), (off[c + 1], off[-1])] for c in range(self.num_classes)]

    def query_state(self, m: int):
        """Filter-side state of m queries: packed bf16 operand, |h|^2, rounding-error norm, running
        minimum bits, candidate count (what tip_nn_query_prep fills)."""
        return (torch.empty((m, self.pitch), dtype=torch.bfloat16, device=self.dev),
                torch.empty(m, dtype=torch.float32, device=self.dev),
                torch.empty(m, dtype=torch.float32, device=self.dev),
                torch.empty(m, dtype=torch.int32, device=self.dev),
                torch.empty(m, dtype=torch.int32, device=self.dev))

    def has_items(self, q_off: np.ndarray, mode: int) -> bool:
        """True if a search of this mode would launch the filter for queries with this class histogram."""
        if self.n == 0:
            return False
        ranges = self.ranges(mode)
        return any(int(q_off[c + 1]) > int(q_off[c]) and any(int(hi) > int(lo) for lo, hi in ranges[c])
                   for c in range(self.num_classes))

    def search(self, q: torch.Tensor, q_class: torch.Tensor, q_off: np.ndarray, mode: int, use_filter: bool = True,
               want_rows: bool = False, prepacked=None, next_query=None, q_idx: Optional[torch.Tensor] = None, fin=None,
               speculative: bool = False):
        """q: [m, d] queries grouped by class (q_off), q_class[m] int32.  Returns, per query, the
        exact NumPy-order distance to its nearest train row in the range selected by `mode`
        (NaN when the range is empty on this shard), that row's position (-1 when empty), its
        original index, and (want_rows) a copy of the winning train rows.
        prepacked: query_state() already filled for q (by a previous search's next_query);
        next_query: query_state() to fill with the winning rows as the next search's queries.
        q_idx: int32 [m] — the queries are rows q_idx[r] of q (class-sorted order over the caller's buffer; no gathered
        copy); fin: (dist_a, gid, idx, n_total, out) — fused result scatter of DSA's last stage (tip_rerank_extras).
        speculative: queries whose candidate list is empty or overflowed are only counted in work_buffer(m)[0] (no
        exhaustive-scan launch); the caller checks overflow_count() and repeats the call without it if non-zero."""
        m = q.shape[0] if q_idx is None else q_idx.shape[0]
        out_dist = torch.empty(m, dtype=q.dtype, device=self.dev)
        out_pos = torch.empty(m, dtype=torch.int32, device=self.dev)
        out_gid = torch.empty(m, dtype=torch.int32, device=self.dev)
        out_rows = torch.empty((m, self.d), dtype=q.dtype, device=self.dev) if want_rows else None
        if m == 0:
            return out_dist, out_pos, out_gid, out_rows
        lib = self.lib
        cand_idx = cand_cnt = None
        if use_filter and self.n > 0:
            ranges = self.ranges(mode)
            key = (mode, np.asarray(q_off, dtype=np.int64).tobytes())
            plan = self._item_cache.get(key)
            if plan is None:       # host planning + upload once per (mode, class histogram)
                populated = int(np.count_nonzero(np.diff(np.asarray(q_off))))
                # few queries per class: tile across class boundaries (full tiles), per-query masking;
                # many: class-aligned tiles never touch their own class's rows
                mixed = mode == _lib.RANGE_OTHER_CLASSES and m < 2 * self.row_tile * max(1, populated)
                if self.resident:
                    # resident-query kernel: one or two long items per CTA, equal tile counts
                    tiles = query_tiles(q_off, ranges, self.class_off, self.row_tile, mixed)
                    items, n_static = build_balanced_items(tiles, self.col_tile, self.sms, pool_frac=POOL_FRAC,
                                                           pool_tiles=POOL_TILES)
                else:
                    n_static = None
                    span_of = lambda pairs: span_tiles_for(pairs, self.sms)
                    if mixed:
                        items = build_other_class_items(q_off, self.class_off, self.row_tile, self.col_tile, span_of)
                    else:
                        pairs = count_tile_pairs(q_off, ranges, self.row_tile, self.col_tile)
                        items, _ = build_items(q_off, ranges, span_of(pairs), self.row_tile, self.col_tile)
                    items = span_major(items)
                flops = 2.0 * self.d * sum(int(q_off[c + 1] - q_off[c]) * sum(int(hi) - int(lo) for lo, hi in ranges[c])
                                           for c in range(self.num_classes))
                flagged = bool(items.shape[0] and (items[:, 5] & 1).any())
                n_static = items.shape[0] if n_static is None else n_static
                plan = (torch.from_numpy(items).to(self.dev) if items.shape[0] else None, items.shape[0], flops, flagged,
                        n_static)
                if len(self._item_cache) > 64:
                    self._item_cache.clear()
                self._item_cache[key] = plan
            items_dev, n_items, flops, flagged, n_static = plan
            if n_items > 0:
                if prepacked is not None:
                    q_pack, q_sq, q_err, row_min, cand_cnt = prepacked
                else:
                    q_pack, q_sq, q_err, row_min, cand_cnt = self.query_state(m)
                    _lib.check(lib.tip_nn_query_prep(_p(q), tip_dtype(q.dtype), m, self.d, _p(self.center), _p(q_pack),
                                                     _p(q_sq), _p(q_err), _p(row_min), _p(cand_cnt), _p(q_idx), _stream()),
                               "tip_nn_query_prep")
                cand_idx = torch.empty((m, self.cap, 2), dtype=torch.int32, device=self.dev)
                ev = None
                if PROFILE is not None:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                _lib.check(lib.tip_nn_filter(_p(q_pack), _p(q_sq), m, _p(self.t_pack), self.n, self.d, self.pitch,
                                             _p(items_dev), n_items, _p(q_class if flagged else None),
                                             _p(self.class_off_dev if flagged else None), self.rmax,
                                             _p(q_err), self.errmax, _p(row_min), _p(cand_idx),
                                             _p(cand_cnt), self.cap, n_static,
                                             _p(self.sched_counter if n_static < n_items else None),
                                             _stream()), "tip_nn_filter")
                if ev is not None:
                    ev[1].record()
                    name = "nn_filter_same_class" if mode == _lib.RANGE_SAME_CLASS else "nn_filter_other_classes"
                    PROFILE.append((name, flops, ev[0], ev[1]))
        work = self.work_buffer(m, q.dtype)
        if self._capture_refs is not None:
            # cached tensors whose addresses get baked into the graph: the plan keeps them alive even
            # if the engine's caches evict them later
            self._capture_refs.extend(t for t in (work, self.sched_counter, self.t_pack, self.t, self.t_gid,
                                                  self.class_off_dev, self.center,
                                                  items_dev if (use_filter and self.n > 0) else None) if t is not None)
        nq = [_p(t) for t in next_query] if next_query is not None else [None] * 5
        # stage 1 -> stage 2 hand-over: the winners' fit-time bounds seed the other-class search's running minima
        seed = self.seed_b if (next_query is not None and mode == _lib.RANGE_SAME_CLASS) else None
        if self._capture_refs is not None and seed is not None:
            self._capture_refs.append(seed)
        ex = _lib.RerankExtras()
        ex.q_idx = 0 if q_idx is None else q_idx.data_ptr()
        ex.next_seed_ub = 0 if seed is None else seed.data_ptr()
        ex.next_t_rmax, ex.next_t_errmax = self.rmax, self.errmax
        ex.count_overflow_only = 1 if (speculative and cand_cnt is not None) else 0
        if fin is not None:
            f_a, f_gid, f_idx, f_n, f_out = fin
            ex.fin_dist_a, ex.fin_gid = f_a.data_ptr(), f_gid.data_ptr()
            ex.fin_idx = 0 if f_idx is None else f_idx.data_ptr()
            ex.fin_n_total, ex.fin_out = int(f_n), f_out.data_ptr()
        _lib.check(lib.tip_nn_rerank(_p(q), _p(self.t), tip_dtype(q.dtype), m, self.n, self.d, _p(cand_idx),
                                     _p(cand_cnt), self.cap, _p(q_class), _p(self.class_off_dev), self.num_classes,
                                     mode, _p(self.t_gid), _p(out_dist), _p(out_pos), _p(out_gid), _p(out_rows),
                                     _p(work), _p(self.stats),
                                     _p(self.center) if next_query is not None else None, *nq,
                                     C.byref(ex), _stream()),
                   "tip_nn_rerank")
        self.last_cand_cnt = cand_cnt
        if cand_cnt is not None:
            self.last_cand_cnt_by_mode[mode] = (cand_cnt, cand_idx)
        return out_dist, out_pos, out_gid, out_rows

    def gather(self, pos: torch.Tensor) -> torch.Tensor:
        out = torch.empty((pos.shape[0], self.d), dtype=self.t.dtype, device=self.dev)
        if pos.shape[0]:
            _lib.check(self.lib.tip_gather_rows(_p(self.t), self.d * self.t.element_size(), _p(pos), pos.shape[0],
                                                _p(out), _stream()), "tip_gather_rows")
        return out


def winner_queries(engine: "NnEngine", p2p: Optional[P2PExchange], gdist: Optional[torch.Tensor],
                   ggid: Optional[torch.Tensor], m: int, dtype: torch.dtype, state2):
    """Global stage-1 winners -> (dist_a, original index, winning rows) + the packed stage-2 query
    state, gathered from the replicated training set (tip_shard_winner_queries).  With `p2p` the
    per-shard records of the exchange started by the last push are reduced inside the kernel."""
    dev = engine.dev
    dist_a = torch.empty(m, dtype=dtype, device=dev)
    gid = torch.empty(m, dtype=torch.int32, device=dev)
    rows = torch.empty((m, engine.d), dtype=dtype, device=dev)
    q_pack, q_sq, q_err, row_min, cand_cnt = state2
    _lib.check(engine.lib.tip_shard_winner_queries(p2p.handle if p2p is not None else None, _p(gdist), _p(ggid),
                                                   tip_dtype(dtype), m, engine.d, _p(engine.t_full),
                                                   engine.t_full.shape[0], _p(engine.center), _p(dist_a), _p(gid), _p(rows),
                                                   _p(q_pack), _p(q_sq), _p(q_err), _p(row_min), _p(cand_cnt), _stream()),
               "tip_shard_winner_queries")
    return dist_a, gid, rows


def dsa_distances(engine: NnEngine, x: torch.Tensor, q_class: torch.Tensor, q_off: np.ndarray,
                  comm: Optional[TrainShardComm] = None, use_filter: bool = True, q_idx: Optional[torch.Tensor] = None,
                  scatter=None, speculative: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """surprise.py:615-631 for class-sorted queries x: (dist_a, dist_b, winner original index).
    Eager launches (the CUDA-graph plans below replay the same sequence).  Single shard only: q_idx — the
    class-sorted queries are rows q_idx[r] of x (the gather is fused into the pack and the re-rank); scatter =
    (idx, n_total, out) — the last re-rank also writes dist_a, dist_b, winner, dist_a / dist_b into out[4, n_total]."""
    sharded = comm is not None and comm.world > 1
    m = x.shape[0] if q_idx is None else q_idx.shape[0]
    if not sharded and engine.table_b is not None and engine.table_b.dtype == x.dtype:
        # fit-time table: stage 1 only, dist_b looked up by the winner's position
        dist_a, pos, gid, _ = engine.search(x, q_class, q_off, _lib.RANGE_SAME_CLASS, use_filter, q_idx=q_idx,
                                            speculative=speculative)
        found = pos >= 0
        dist_b = torch.where(found, engine.table_b.index_select(0, pos.clamp_min(0).to(torch.int64)),
                             torch.full_like(dist_a, float("nan")))
        if scatter is not None:
            _lib.check(engine.lib.tip_dsa_pack_out(_p(dist_a), _p(dist_b), tip_dtype(dist_a.dtype), _p(gid), _p(scatter[0]),
                                                   m, int(scatter[1]), _p(scatter[2]), _stream()), "tip_dsa_pack_out")
        return dist_a, dist_b, gid
    if not sharded:
        # single shard: stage 1's re-rank also emits its winners as the packed queries of stage 2
        fuse = use_filter and engine.has_items(q_off, _lib.RANGE_OTHER_CLASSES)
        state2 = engine.query_state(m) if fuse else None
        dist_a, _, gid, winners = engine.search(x, q_class, q_off, _lib.RANGE_SAME_CLASS, use_filter, want_rows=True,
                                                next_query=state2, q_idx=q_idx, speculative=speculative)
        fin = None if scatter is None else (dist_a, gid, scatter[0], scatter[1], scatter[2])
        dist_b = engine.search(winners, q_class, q_off, _lib.RANGE_OTHER_CLASSES, use_filter, prepacked=state2, fin=fin,
                               speculative=speculative)[0]
        return dist_a, dist_b, gid
    assert q_idx is None and scatter is None
    replica = engine.t_full is not None
    p2p = comm.p2p(engine.dev, m) if replica else None
    local_a, _, local_gid, local_rows = engine.search(x, q_class, q_off, _lib.RANGE_SAME_CLASS, use_filter,
                                                      want_rows=not replica)
    if replica:
        state2 = engine.query_state(m)
        if p2p is not None:
            p2p.push_nn(local_a, local_gid)
            dist_a, gid, winners = winner_queries(engine, p2p, None, None, m, x.dtype, state2)
        else:
            gmin, ggid = comm.reduce_winner_index(local_a, local_gid)
            dist_a, gid, winners = winner_queries(engine, None, gmin.contiguous(), ggid.to(torch.int32).contiguous(), m,
                                                  x.dtype, state2)
        prepacked = state2 if (use_filter and engine.has_items(q_off, _lib.RANGE_OTHER_CLASSES)) else None
    else:
        dist_a, gid, winners = comm.reduce_winners(local_a, local_gid, local_rows)
        prepacked = None
    local_b = engine.search(winners, q_class, q_off, _lib.RANGE_OTHER_CLASSES, use_filter, prepacked=prepacked)[0]
    if p2p is not None:
        p2p.push_nn(local_b, None)
        dist_b = torch.empty_like(local_b)
        p2p.min_into(dist_b)
    else:
        dist_b = comm.reduce_min_nan(local_b)
    return dist_a, dist_b, gid


class _Capture:
    """Context for capturing engine work into a CUDA graph: garbage from earlier engines (old graphs,
    their private pools) must not be released in the middle of the capture — a cudaFree / graph
    destroy there invalidates it — and engine-owned tensors whose addresses get baked into the
    graph are kept alive by the plan."""

    def __init__(self, engine: "NnEngine", keep: list):
        self.engine, self.keep = engine, keep

    def __enter__(self):
        import gc

        gc.collect()
        self.was_enabled = gc.isenabled()
        gc.disable()
        self.engine._capture_refs = []
        return self

    def __exit__(self, *exc):
        import gc

        self.keep.extend(self.engine._capture_refs)
        self.engine._capture_refs = None
        if self.was_enabled:
            gc.enable()
        return False


class DsaPlan:
    """The whole scoring call for one (batch size, class histogram) captured as CUDA graph(s):
    gather into class-sorted order -> pack -> filter -> re-rank(+winner rows) -> pack -> filter ->
    re-rank -> scatter of (dist_a, dist_b, winner index) back to the caller's row order.
    Inputs: `x_in` [n_total, d] (the engine's landing buffer for host uploads) and `idx` [m]
    (original row of every class-sorted query).  Output: `out` [4, n_total] float64 in the
    caller's order — dist_a, dist_b, winner index, dist_a / dist_b (divided in the trace dtype) —
    where rows the reference never scores keep NaN / -1.

    N_train sharded (comm.world > 1): with the peer-memory exchange the two exchange steps are
    kernels of ours and the call is still ONE graph; with the torch.distributed fallback the call is
    three graph segments with one eager MIN all-reduce between consecutive segments (two for float64
    traces in the first exchange)."""

    def __init__(self, engine: "NnEngine", m: int, q_off: np.ndarray, dtype: torch.dtype, use_filter: bool,
                 comm: Optional[TrainShardComm] = None, n_total: Optional[int] = None):
        self.engine = engine
        self.q_off = np.asarray(q_off, dtype=np.int64).copy()
        dev = engine.dev
        lib = engine.lib
        self.m = int(m)
        self.n_total = int(m if n_total is None else n_total)
        self.x_in = engine.input_buffer(self.n_total, dtype)
        self.idx = torch.arange(self.m, dtype=torch.int32, device=dev)
        self.x = torch.zeros((m, engine.d), dtype=dtype, device=dev)
        self.out = torch.full((4, self.n_total), float("nan"), dtype=torch.float64, device=dev)
        self.out[2].fill_(-1.0)
        self.out_host = torch.empty((4, self.n_total), dtype=torch.float64).pin_memory()
        self.idx_host = torch.empty(self.m, dtype=torch.int32).pin_memory()
        q_class = np.repeat(np.arange(engine.num_classes, dtype=np.int32), np.diff(self.q_off))
        self.q_class = torch.from_numpy(q_class).to(dev)
        row_bytes = engine.d * self.x.element_size()
        self.comm = comm if (comm is not None and comm.world > 1) else None
        sharded = self.comm is not None
        if sharded and engine.t_full is None:
            raise RuntimeError("the sharded plan needs the replicated training set (NnEngine.t_full)")
        self.p2p = self.comm.p2p(dev, self.m) if sharded else None
        # Single shard with the filter on: the replayed graph has NO exhaustive-scan launches (two normally empty
        # launches per call); queries whose candidate list is empty or overflowed are only counted, and the caller
        # repeats such a call on the eager path (overflowed() / clear_overflow()).
        self.speculative = bool(use_filter and not sharded and SPECULATIVE)
        self.overflow = engine.work_buffer(self.m, dtype)[:4].view(torch.int32)
        self.overflow_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._keep_alive = []
        S = self.state = {}

        def gather():
            _lib.check(lib.tip_gather_rows(_p(self.x_in), row_bytes, _p(self.idx), self.m, _p(self.x), _stream()),
                       "tip_gather_rows")

        def scatter(a, b, gid):
            if self.n_total != self.m:         # which rows are unscored can change between calls
                self.out.fill_(float("nan"))
                self.out[2].fill_(-1.0)
            _lib.check(lib.tip_dsa_pack_out(_p(a), _p(b), tip_dtype(a.dtype), _p(gid), _p(self.idx), self.m,
                                            self.n_total, _p(self.out), _stream()), "tip_dsa_pack_out")
            self.dist_a, self.dist_b, self.gid = a, b, gid

        if not sharded:
            def whole():
                # gather fused into the pack / re-rank (q_idx), result scatter fused into the last re-rank
                if self.n_total != self.m:         # which rows are unscored can change between calls
                    self.out.fill_(float("nan"))
                    self.out[2].fill_(-1.0)
                self.dist_a, self.dist_b, self.gid = dsa_distances(engine, self.x_in, self.q_class, self.q_off, None, use_filter,
                                                                   q_idx=self.idx, scatter=(self.idx, self.n_total, self.out),
                                                                   speculative=self.speculative)

            segments = [(whole, True)]
        elif self.p2p is not None:
            def whole():
                gather()
                scatter(*dsa_distances(engine, self.x, self.q_class, self.q_off, self.comm, use_filter))

            segments = [(whole, True)]
        else:
            # torch.distributed fallback: static exchange buffers written by one graph segment,
            # all-reduced in place (eager), read by the next segment
            f32 = dtype == torch.float32
            S["keys"] = torch.zeros(self.m, dtype=torch.int64 if f32 else dtype, device=dev)
            S["cand"] = torch.zeros(self.m, dtype=torch.int64, device=dev)
            S["bmin"] = torch.zeros(self.m, dtype=dtype, device=dev)
            big = torch.iinfo(torch.int64).max

            def seg_a():
                gather()
                la, _, lg, _ = engine.search(self.x, self.q_class, self.q_off, _lib.RANGE_SAME_CLASS, use_filter)
                S["la"], S["lg"] = la, lg
                if f32:
                    S["keys"].copy_(pack_winner_keys(la, lg))
                else:
                    S["keys"].copy_(torch.where(torch.isnan(la), torch.full_like(la, float("inf")), la))

            def exch_1():
                self.comm.min_(S["keys"])
                if not f32:
                    la = torch.where(torch.isnan(S["la"]), torch.full_like(S["la"], float("inf")), S["la"])
                    S["cand"].copy_(torch.where((la == S["keys"]) & (S["lg"] >= 0), S["lg"].to(torch.int64),
                                                torch.full_like(S["cand"], big)))
                    self.comm.min_(S["cand"])

            def seg_b():
                if f32:
                    gmin, ggid = unpack_winner_keys(S["keys"])
                else:
                    none = torch.isinf(S["keys"]) | (S["cand"] == big)
                    gmin = torch.where(none, torch.full_like(S["keys"], float("nan")), S["keys"])
                    ggid = torch.where(none, torch.full_like(S["cand"], -1), S["cand"])
                state2 = engine.query_state(self.m)
                a, gid, winners = winner_queries(engine, None, gmin.contiguous(), ggid.to(torch.int32).contiguous(),
                                                 self.m, dtype, state2)
                S["a"], S["gid"] = a, gid
                pre = state2 if (use_filter and engine.has_items(self.q_off, _lib.RANGE_OTHER_CLASSES)) else None
                lb = engine.search(winners, self.q_class, self.q_off, _lib.RANGE_OTHER_CLASSES, use_filter, prepacked=pre)[0]
                S["bmin"].copy_(torch.where(torch.isnan(lb), torch.full_like(lb, float("inf")), lb))

            def exch_2():
                self.comm.min_(S["bmin"])

            def seg_c():
                b = torch.where(torch.isinf(S["bmin"]), torch.full_like(S["bmin"], float("nan")), S["bmin"])
                scatter(S["a"], b, S["gid"])

            segments = [(seg_a, True), (exch_1, False), (seg_b, True), (exch_2, False), (seg_c, True)]

        # eager warm-up on a side stream: fills caches, sets kernel attributes, runs the collectives once
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for fn, _ in segments:
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.steps = []
        for fn, capturable in segments:
            if capturable:
                g = torch.cuda.CUDAGraph()
                with _Capture(engine, self._keep_alive):
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        fn()
                self.steps.append(g.replay)
            else:
                fn()                      # keeps the ranks' collective sequences aligned during set-up
                self.steps.append(fn)
        self.graph = None
        if len(segments) == 1:
            self.graph = g
        if self.speculative:
            self.overflow.zero_()     # whatever the warm-up ran on (a stale landing buffer) does not count

    def load_sorted(self, x_sorted: torch.Tensor):
        """Device-resident input already in class-sorted order (benchmarks, tools)."""
        assert self.n_total == self.m
        self.x_in.copy_(x_sorted)
        self.idx.copy_(torch.arange(self.m, dtype=torch.int32, device=self.idx.device))

    def run(self):
        for step in self.steps:
            step()
        return self.out

    def fetch_overflow(self):
        """enqueue the D2H copy of the overflow counter (speculative plans); read it with overflowed() after a sync"""
        if self.speculative:
            self.overflow_host.copy_(self.overflow, non_blocking=True)

    def overflowed(self) -> bool:
        return self.speculative and int(self.overflow_host[0]) != 0

    def clear_overflow(self):
        self.overflow.zero_()
        self.overflow_host.zero_()


def dsa_plan(engine: "NnEngine", m: int, q_off: np.ndarray, dtype: torch.dtype, use_filter: bool,
             comm: Optional[TrainShardComm] = None, n_total: Optional[int] = None) -> DsaPlan:
    n_total = int(m if n_total is None else n_total)
    key = (m, n_total, np.asarray(q_off, dtype=np.int64).tobytes(), dtype, use_filter, engine.cap, comm is not None,
           engine.table_b is not None)
    plan = engine._plans.get(key)
    if plan is None:
        if len(engine._plans) >= 8:
            engine._plans.pop(next(iter(engine._plans)))
        plan = DsaPlan(engine, m, q_off, dtype, use_filter, comm, n_total=n_total)
        engine._plans[key] = plan
    return plan


# ------------------------------------------------------------------------------------------
# Gaussian-KDE engine (LSA)
# ------------------------------------------------------------------------------------------
F16_NORM_SCALE = 64.0       # tail scale of the fp16 operand: |p|^2 / 2 / 64 must stay below 65504


class KdeEngine:
    def __init__(self, p_whitened: np.ndarray, comm: Optional[TrainShardComm] = None):
        """p_whitened: [n, d] float64 whitened, centred training traces (host).  With a communicator
        of more than one rank the engine keeps rows rank::world (N_train sharded) and merges the
        per-shard partial sums in log_kernel_sum (north_star: one exchange of partial KDE sums).

        Two packed operands are kept: the three-segment split-bf16 form (~2^-17 relative per product: always
        accurate enough for rtol 1e-4) and a one-segment fp16 form (11-bit significands, a third of the tensor
        work).  Which one a scoring call uses is decided by MEASURING the fast form's error on a sample of the
        queries (core/stable_kde.py)."""
        self.dev = require_cuda()
        self.lib = _lib.load()
        self.comm = comm if (comm is not None and comm.world > 1) else None
        if self.comm is not None:
            p_whitened = np.ascontiguousarray(p_whitened[self.comm.rank::self.comm.world])
        self.n, self.d = p_whitened.shape
        self.precision = "split-bf16 x3 (h.h + h.l + l.h, ~2^-17 relative)"
        self.pitch = int(self.lib.tip_pair_pitch(self.d, 3))
        self.pitch1 = int(self.lib.tip_pair_pitch(self.d, 1))
        sms = C.c_int(0)
        _lib.check(self.lib.tip_device_info(C.byref(sms), None, None), "tip_device_info")
        self.sms = sms.value
        p32 = to_device(p_whitened.astype(np.float32), self.dev)
        self.t_pack = torch.empty((self.n, self.pitch), dtype=torch.bfloat16, device=self.dev)
        _lib.check(self.lib.tip_pair_prep(_p(p32), _lib.TIP_F32, self.n, self.d, None, _lib.ROLE_TRAIN, 3, 1.0, -0.5,
                                          _p(self.t_pack), None, None, _stream()), "tip_pair_prep")
        # fast form; unusable if a training value or its norm term leaves fp16's range
        self.t_pack_f16 = torch.empty((self.n, self.pitch1), dtype=torch.float16, device=self.dev)
        self.flags = torch.zeros(1, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.tip_pair_prep_f16(_p(p32), _lib.TIP_F32, self.n, self.d, None, _lib.ROLE_TRAIN, -0.5,
                                              F16_NORM_SCALE, _p(self.t_pack_f16), None, _p(self.flags), _stream()),
                   "tip_pair_prep_f16")
        torch.cuda.current_stream().synchronize()
        self.fast_ok = (os.environ.get("B200TIP_LSA_FAST", "1") != "0" and int(self.flags.item()) == 0
                        and self.comm is None)
        if not self.fast_ok:
            self.t_pack_f16 = None
        self._items = {}

    def _work_items(self, m: int):
        """Host planning + upload once per batch size."""
        plan = self._items.get(m)
        if plan is None:
            q_off = np.array([0, m], dtype=np.int64)
            ranges = [[(0, self.n)]]
            rt = int(self.lib.tip_kde_tile_rows())
            items, slots = build_items(q_off, ranges, span_tiles_for(count_tile_pairs(q_off, ranges, rt), self.sms), rt)
            items = span_major(items)
            if len(self._items) > 16:
                self._items.clear()
            # one partial per column part of a tile (tip_kde_slot_parts: 4 x 64 columns)
            plan = (torch.from_numpy(items).to(self.dev), items.shape[0], slots * int(self.lib.tip_kde_slot_parts()))
            self._items[m] = plan
        return plan

    def log_kernel_sum(self, q: torch.Tensor, fast: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """q: [m, d] fp32 whitened, centred queries on the GPU.  Returns (mx, sm, qsq) with
        sum_i exp(-|p_i - q_j|^2/2) = exp(mx_j - qsq_j/2) * sm_j.  fast=True: the one-segment fp16 pass
        (sets flags[0] if a query leaves fp16's range)."""
        m = q.shape[0]
        lib = self.lib
        q_sq = torch.empty(m, dtype=torch.float32, device=self.dev)
        if fast:
            assert self.fast_ok
            q_pack = torch.empty((m, self.pitch1), dtype=torch.float16, device=self.dev)
            _lib.check(lib.tip_pair_prep_f16(_p(q), _lib.TIP_F32, m, self.d, None, _lib.ROLE_QUERY, 0.0, F16_NORM_SCALE,
                                             _p(q_pack), _p(q_sq), _p(self.flags), _stream()), "tip_pair_prep_f16")
        else:
            q_pack = torch.empty((m, self.pitch), dtype=torch.bfloat16, device=self.dev)
            _lib.check(lib.tip_pair_prep(_p(q), _lib.TIP_F32, m, self.d, None, _lib.ROLE_QUERY, 3, 1.0, 0.0, _p(q_pack),
                                         _p(q_sq), None, _stream()), "tip_pair_prep")
        items_dev, n_items, slots = self._work_items(m)
        part_max = torch.full((slots, m), float("-inf"), dtype=torch.float32, device=self.dev)
        part_sum = torch.zeros((slots, m), dtype=torch.float32, device=self.dev)
        if fast:
            _lib.check(lib.tip_kde_lse_f16(_p(q_pack), m, _p(self.t_pack_f16), self.n, self.d, self.pitch1, _p(items_dev),
                                           n_items, _p(part_max), _p(part_sum), _stream()), "tip_kde_lse_f16")
        else:
            _lib.check(lib.tip_kde_lse(_p(q_pack), m, _p(self.t_pack), self.n, self.d, self.pitch, _p(items_dev),
                                       n_items, _p(part_max), _p(part_sum), _stream()), "tip_kde_lse")
        mx = torch.empty(m, dtype=torch.float32, device=self.dev)
        sm = torch.empty(m, dtype=torch.float32, device=self.dev)
        _lib.check(lib.tip_kde_combine(_p(part_max), _p(part_sum), m, slots, _p(mx), _p(sm), _stream()),
                   "tip_kde_combine")
        if self.comm is not None:
            p2p = self.comm.p2p(self.dev, m)
            mx, sm = p2p.lse(mx, sm) if p2p is not None else self.comm.reduce_lse(mx, sm)
        return mx, sm, q_sq


def whiten(x: torch.Tensor, cols: Optional[torch.Tensor], mu: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """(x[:, cols] - mu) @ w in fp32 on the GPU (scipy 1.4.1: points . cholesky(inv_cov))."""
    lib = _lib.load()
    m, d_in = x.shape
    d_out = w.shape[0]
    out = torch.empty((m, d_out), dtype=torch.float32, device=x.device)
    _lib.check(lib.tip_whiten(_p(x), tip_dtype(x.dtype), m, d_in, _p(cols), d_out, _p(mu), _p(w), _p(out), _stream()),
               "tip_whiten")
    return out


def quadratic_forms(x, centres: Sequence[Tuple[torch.Tensor, torch.Tensor]], cols: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q[k, j] = |(x_j - mu_k) . W_k|^2 for every (mu_k [d] float64, W_k [d, d] float32) pair: whitening on the
    GPU (tip_whiten, centred first) + a double-accumulated row norm (tip_row_sqnorm).  x: [m, d_in] device
    matrix (float32 / float64).  Returns float64 [len(centres), m] on the device."""
    lib = _lib.load()
    m = x.shape[0]
    out = torch.empty((len(centres), m), dtype=torch.float64, device=x.device)
    for k, (mu, w) in enumerate(centres):
        y = whiten(x, cols, mu, w)
        if m:
            _lib.check(lib.tip_row_sqnorm(_p(y), m, y.shape[1], _p(out[k]), _stream()), "tip_row_sqnorm")
    return out
