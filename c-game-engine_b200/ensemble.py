"""Multi-GPU host logic for ensembles of independent worlds.

A single world does not shard (one island graph, Gauss-Seidel coupling), so the scaling axis is the number of world
instances: rank r of R owns the contiguous block of world indices [first, first + count).  Nothing is exchanged
inside the tick.  The only collective is the end-of-run gather of the 32-byte per-world stats record
(gpx_world_stats) — NCCL on the GPU box, gloo in the CPU tests — plus a MAX reduction of the device time.
"""
from __future__ import annotations

import numpy as np

STATS_DTYPE = np.dtype([("kinetic_energy", "<f4"), ("max_speed", "<f4"), ("awake_bodies", "<u4"),
                        ("manifolds", "<u4"), ("position_checksum", "<u8"), ("ticks", "<u4"), ("error", "<u4")])

_FNV_OFFSET = 1469598103934665603
_FNV_PRIME = 1099511628211
_M64 = (1 << 64) - 1


def shard(total_worlds: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of world indices owned by `rank`: worlds [g*T/G, (g+1)*T/G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    first = total_worlds * rank // world_size
    last = total_worlds * (rank + 1) // world_size
    return first, last - first


def position_checksum(xf: np.ndarray) -> int:
    """Order-independent checksum of the exact position/rotation bits of one world, as k_stats computes it:
    sum over bodies i of FNV-1a over the 7 words seeded with offset * (i + 1)."""
    total = 0
    words = np.ascontiguousarray(xf, np.float32).view(np.uint32).reshape(-1, 7)
    for i, row in enumerate(words):
        h = (_FNV_OFFSET * (i + 1)) & _M64
        for w in row:
            h = ((h ^ int(w)) * _FNV_PRIME) & _M64
        total = (total + h) & _M64
    return total


def host_stats(xf: np.ndarray, vel: np.ndarray, mass: float, ticks: int, manifolds: int = 0, error: int = 0):
    """The gpx_world_stats record of one world computed on the host from its transforms (n,7) and velocities (n,6)."""
    s = np.zeros((), STATS_DTYPE)
    v2 = (vel[:, :3].astype(np.float32) ** 2).sum(axis=1)
    s["kinetic_energy"] = np.float32(0.5 * mass) * v2.sum(dtype=np.float32)
    s["max_speed"] = np.sqrt(v2.max()) if len(v2) else 0.0
    s["awake_bodies"] = len(xf)
    s["manifolds"] = manifolds
    s["position_checksum"] = position_checksum(xf)
    s["ticks"] = ticks
    s["error"] = error
    return s


def gather_stats(stats: np.ndarray, dist=None, device: str = "cpu") -> np.ndarray:
    """All ranks' per-world stats concatenated in rank (= world index) order.  `dist` is torch.distributed or None."""
    stats = np.ascontiguousarray(stats, STATS_DTYPE)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return stats
    import torch
    ws = dist.get_world_size()
    n = torch.tensor([len(stats)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    buf = np.zeros(cap * STATS_DTYPE.itemsize, np.uint8)
    buf[:len(stats) * STATS_DTYPE.itemsize] = stats.view(np.uint8).reshape(-1)
    t = torch.from_numpy(buf).to(device)
    out = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(out, t)
    parts = [o.cpu().numpy()[:c * STATS_DTYPE.itemsize].view(STATS_DTYPE) for o, c in zip(out, counts)]
    return np.concatenate(parts)


def max_over_ranks(x: float, dist=None, device: str = "cuda") -> float:
    """Multi-GPU timings are reported as the max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
