"""TimeLoop-shaped step drivers over the CUDA engines.

The reference consumes its back-ends through earth2mip's ``TimeLoop`` protocol
(/root/reference/skyrim/core/models/utils.py:20,25-29,34,45-47):
``model(time, x) -> Iterator[(time, Tensor[B,C,H,W], restart)]`` whose FIRST yield is the
initial condition itself, plus the attributes ``n_history_levels``, ``device``,
``in_channel_names``, ``out_channel_names``, ``grid.lat``, ``grid.lon``, ``time_step``.
These classes keep that protocol so ``run_basic_inference`` can drive them unchanged, with the
state resident on the GPU between steps (the reference round-trips it through the host every
step, utils.py:24-31,37).
"""
from __future__ import annotations

import datetime
from dataclasses import dataclass

import numpy as np

from .config import FCNV2_CHANNELS, PANGU_CHANNELS


@dataclass
class Grid:
    lat: list
    lon: list

    @property
    def shape(self):
        return (len(self.lat), len(self.lon))


def equiangular_grid(nlat: int = 721, nlon: int = 1440) -> Grid:
    # pangu.py:33-36 docstring: lat 90 .. -90 (721), lon 0 .. 359.75 (1440)
    return Grid(lat=np.linspace(90.0, -90.0, nlat).tolist(), lon=(np.arange(nlon) * (360.0 / nlon)).tolist())


class _EngineTimeLoop:
    n_history_levels = 1
    time_step = datetime.timedelta(hours=6)
    channel_names: list = []

    def __init__(self, engine):
        import torch
        self.torch = torch
        self.engine = engine
        self.device = torch.device("cuda", engine.device)
        self.in_channel_names = list(self.channel_names)
        self.out_channel_names = list(self.channel_names)
        self.grid = equiangular_grid(engine.cfg.nlat, engine.cfg.nlon)
        self._dev_in = None
        self._dev_out = None
        self._host_out = None

    # earth2mip TimeLoops accept .to(device); ours is pinned to its GPU (ensemble.py:34,46 calls these)
    def to(self, device):
        return self

    def cuda(self):
        return self

    def __call__(self, time, x, restart=None):
        """x: (B, n_history_levels, C, H, W) tensor (host or device).  Yields (time, state, None):
        first the initial condition, then one 6-h step per iteration, state resident in HBM."""
        torch = self.torch
        assert x.dim() == 5 and x.shape[1] == self.n_history_levels, x.shape
        cur = x[:, -1].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        yield time, cur.clone(), None
        nxt = torch.empty_like(cur)
        while True:
            self.engine.step(cur, nxt)
            time = time + self.time_step
            yield time, nxt, None
            cur, nxt = nxt, torch.empty_like(cur)  # the yielded tensor stays valid for the caller

    def step_host(self, x_host):
        """One step with HOST (pinned) input and HOST (pinned) output — the end-to-end unit that
        bench.py's ``e2e`` times: H2D of the state, sky_model_step, D2H of the result."""
        torch = self.torch
        if self._dev_in is None or self._dev_in.shape != x_host.shape:
            self._dev_in = torch.empty(x_host.shape, dtype=torch.float32, device=self.device)
            self._dev_out = torch.empty_like(self._dev_in)
            self._host_out = [torch.empty(x_host.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._flip = 0
        self._dev_in.copy_(x_host, non_blocking=True)
        self.engine.step(self._dev_in, self._dev_out)
        out = self._host_out[self._flip]
        self._flip ^= 1
        out.copy_(self._dev_out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out


class PanguTimeLoop(_EngineTimeLoop):
    channel_names = PANGU_CHANNELS


class SFNOTimeLoop(_EngineTimeLoop):
    channel_names = FCNV2_CHANNELS
