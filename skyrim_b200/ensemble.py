"""Perturbed-initial-condition ensembles sharded over GPUs (new functionality —
the reference only has a sequential multi-model mean, models/ensemble.py:86-101).

One process per GPU.  Members are independent trajectories: member m lives on rank m // M
(M members per GPU, stacked along the batch of every kernel so weights and tables are read once
per step).  The only collective of the whole run is ONE broadcast of the fp32 weight arena at
init; optional mean / spread reductions happen after the step loop.
"""
from __future__ import annotations

import numpy as np


def member_range(rank: int, members_per_gpu: int):
    """global member indices owned by ``rank``"""
    return range(rank * members_per_gpu, (rank + 1) * members_per_gpu)


def broadcast_arena(weights, shapes, device, rank: int, world: int):
    """Rank 0 packs the named fp32 tensors; every rank returns (arena tensor on ``device``,
    manifest).  ``weights`` may be None on ranks != 0; ``shapes`` (name -> shape) must be known
    everywhere (it follows from the config)."""
    import torch
    import torch.distributed as dist
    from .engine import pack_arena
    if rank == 0:
        arena_h, manifest = pack_arena(weights)
        arena = torch.from_numpy(arena_h).to(device)
    else:
        arena_h, manifest = pack_arena({k: np.zeros(s, np.float32) for k, s in shapes.items()})
        arena = torch.empty(arena_h.size, dtype=torch.float32, device=device)
    if world > 1:
        dist.broadcast(arena, 0)
    return arena, manifest


class EnsembleRunner:
    """Per-rank driver: M members resident on this GPU, chained device-resident steps."""

    def __init__(self, engine, base_state, sigma_c, members_per_gpu: int, rank: int = 0, amp: float = 0.05,
                 seed: int = 0):
        import torch
        from .engine import perturb_ic
        self.engine, self.M, self.rank = engine, members_per_gpu, rank
        dev = torch.device("cuda", engine.device)
        x = torch.as_tensor(base_state, dtype=torch.float32)[None].repeat(self.M, 1, 1, 1).to(dev).contiguous()
        self.sigma = torch.as_tensor(sigma_c, dtype=torch.float32, device=dev)
        perturb_ic(x, self.sigma, amp, seed=seed, member0=rank * self.M)  # Philox keyed by the GLOBAL member id
        self.x, self.y = x, torch.empty_like(x)
        self.steps = 0

    def step(self, n: int = 1):
        for _ in range(n):
            self.engine.step(self.x, self.y)
            self.x, self.y = self.y, self.x
        self.steps += n
        return self.x

    def mean_and_spread(self, world: int = 1):
        """ensemble mean / standard deviation over ALL members (two all-reduces, outside the step loop)"""
        return mean_and_spread(self.x, world)


def mean_and_spread(x, world: int = 1):
    """x: this rank's members (M, C, H, W).  Two-pass: the mean is reduced first, then the sum of squared DEVIATIONS —
    E[x^2] - mean^2 in fp32 cancels catastrophically on de-normalised fields (z ~ 5e4, t ~ 280 with a spread of a few
    units), and clamping hid the negative variances it produced."""
    import torch
    import torch.distributed as dist
    n = world * x.shape[0]
    s = x.sum(0, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(s)
    mean = (s / n).to(x.dtype)
    ss = ((x - mean) ** 2).sum(0, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ss)
    return mean, (ss / n).sqrt().to(x.dtype)
