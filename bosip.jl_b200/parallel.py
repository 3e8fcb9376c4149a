"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The path shards by contiguous ranges of query/candidate points (global indices preserved); the fit state is tiny and is
replicated by conditioning the same (X, Y, hyper-parameters) on every rank, or broadcast from rank 0.  The only
exchange steps are the ones SURVEY.md section 8(e) lists:
  * acquisition argmax: all-gather of per-rank (value, global index, x) + identical local reduce, ties -> lowest index
  * evidence / MMD sums: all-reduce(sum) of 1-3 doubles
  * TV: all-gather of per-rank (max, sum) log-sum-exp pairs, combined in rank order
Everything degrades to the identity when torch.distributed is not initialised (single GPU)."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def rank_world() -> Tuple[int, int]:
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_range(total: int, rank: int | None = None, world: int | None = None) -> Tuple[int, int]:
    """Contiguous shard (start, count) of `total` items; the first total % world ranks get one extra item."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(int(total), world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def _device():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _allgather_f64(vec: np.ndarray) -> np.ndarray:
    """all-gather a small float64 vector -> (world, len) array; identity for a single process."""
    d = _dist()
    if d is None:
        return np.asarray(vec, dtype=np.float64)[None, :]
    import torch
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=_device())
    outs = [torch.empty_like(t) for _ in range(d.get_world_size())]
    d.all_gather(outs, t)
    return torch.stack(outs).cpu().numpy()


def combine_argmax(vals: np.ndarray, idxs: np.ndarray) -> int:
    """Index (into the gathered arrays) of the winner with Julia's `argmax` semantics: largest value, lowest global index on
    ties, -Inf an ordinary value, NaN above everything (first NaN wins); idx < 0 = that rank had no candidate."""
    best = -1
    for r in range(len(vals)):
        if idxs[r] < 0:
            continue
        if best < 0:
            best = r
            continue
        a_nan, b_nan = np.isnan(vals[best]), np.isnan(vals[r])
        if a_nan or b_nan:
            if b_nan and (not a_nan or idxs[r] < idxs[best]):
                best = r
        elif vals[r] > vals[best] or (vals[r] == vals[best] and idxs[r] < idxs[best]):
            best = r
    return best


def argmax_allgather(idx: int, val: float, x: np.ndarray):
    """Combine per-rank (global index, value, x) winners; every rank returns the same global winner."""
    x = np.asarray(x, dtype=np.float64)
    # the index travels as two exactly-representable halves so that indices >= 2^53 survive the float64 transport
    lo, hi = float(idx & 0xFFFFFFFF) if idx >= 0 else -1.0, float(idx >> 32) if idx >= 0 else -1.0
    g = _allgather_f64(np.concatenate([[val, lo, hi], x]))
    idxs = np.array([(-1 if row[1] < 0 else (int(row[2]) << 32) | int(row[1])) for row in g], dtype=np.int64)
    w = combine_argmax(g[:, 0], idxs)
    if w < 0:
        return -1, -math.inf, x
    return int(idxs[w]), float(g[w, 0]), g[w, 3:].copy()


def _pack_winner(idx: int, val: float, x: np.ndarray) -> np.ndarray:
    lo, hi = (float(idx & 0xFFFFFFFF), float(idx >> 32)) if idx >= 0 else (-1.0, -1.0)
    return np.concatenate([[val, lo, hi], np.asarray(x, dtype=np.float64)])


def _unpack_winners(g: np.ndarray):
    idxs = np.array([(-1 if row[1] < 0 else (int(row[2]) << 32) | int(row[1])) for row in g], dtype=np.int64)
    w = combine_argmax(g[:, 0], idxs)
    if w < 0:
        return -1, -math.inf, g[0, 3:].copy()
    return int(idxs[w]), float(g[w, 0]), g[w, 3:].copy()


class DeviceArgmaxGather:
    """Device-resident argmax exchange: each step's per-rank winner (value, global index, x) is uploaded from a pinned staging
    slot and all-gathered on the device with NO host synchronisation; `result()` reads the last gathered table once (after the
    caller's own synchronise) and reduces it with `combine_argmax`.  Single process: keeps the local winner."""

    def __init__(self, d: int, slots: int = 64):
        import torch
        self.dist = _dist()
        self.d, self.slots, self.k = d, slots, 0
        self.local = None
        if self.dist is not None:
            self.world = self.dist.get_world_size()
            dev = _device()
            pin = dev.type == "cuda"
            self.stage = torch.empty((slots, 3 + d), dtype=torch.float64, pin_memory=pin)
            self.send = torch.empty(3 + d, dtype=torch.float64, device=dev)
            self.table = torch.empty((self.world, 3 + d), dtype=torch.float64, device=dev)

    def push(self, idx: int, val: float, x=None):
        x = np.zeros(self.d) if x is None else x
        if self.dist is None:
            self.local = (int(idx), float(val), np.asarray(x, dtype=np.float64))
            return
        import torch
        slot = self.stage[self.k % self.slots]
        slot.copy_(torch.from_numpy(_pack_winner(idx, val, x)))
        self.k += 1
        self.send.copy_(slot, non_blocking=True)
        self.dist.all_gather_into_tensor(self.table.view(-1), self.send)

    def result(self):
        if self.dist is None:
            return self.local if self.local is not None else (-1, -math.inf, np.zeros(self.d))
        return _unpack_winners(self.table.cpu().numpy())


def allreduce_sum(v: np.ndarray) -> np.ndarray:
    """Sum of a small float64 vector over ranks, accumulated in RANK ORDER on every rank (deterministic, identical)."""
    g = _allgather_f64(np.asarray(v, dtype=np.float64))
    out = np.zeros(g.shape[1])
    for r in range(g.shape[0]):
        out += g[r]
    return out


def allreduce_count(n: int) -> int:
    g = _allgather_f64(np.array([float(n)]))
    return int(round(g[:, 0].sum()))


def lse_merge(m1: float, s1: float, m2: float, s2: float) -> Tuple[float, float]:
    if m2 == -math.inf:
        return m1, s1
    if m1 == -math.inf:
        return m2, s2
    if m1 >= m2:
        return m1, (s1 if m1 == math.inf else s1 + s2 * math.exp(m2 - m1))
    return m2, (s2 if m2 == math.inf else s2 + s1 * math.exp(m1 - m2))


def lse_allgather(m: float, s: float) -> Tuple[float, float]:
    """Combine per-rank (max, sum exp(. - max)) pairs in rank order."""
    g = _allgather_f64(np.array([m, s]))
    M, S = -math.inf, 0.0
    for r in range(g.shape[0]):
        M, S = lse_merge(M, S, float(g[r, 0]), float(g[r, 1]))
    return M, S


def broadcast_fit_inputs(arrays: dict, src: int = 0) -> dict:
    """Broadcast the (small) fit inputs X, Y, hyper-parameters from rank `src` so that every rank conditions the same GP.
    ('factorise on rank 0 or redundantly on all -- N^3/3 is tiny'; redundant conditioning of identical
    inputs gives bit-identical factors on identical GPUs and needs no 134 MB L^-1 broadcast.)"""
    d = _dist()
    if d is None:
        return arrays
    import torch
    out = {}
    for k in sorted(arrays):
        a = np.ascontiguousarray(arrays[k], dtype=np.float64)
        t = torch.as_tensor(a, device=_device())
        d.broadcast(t, src=src)
        out[k] = t.cpu().numpy().reshape(a.shape)
    return out
