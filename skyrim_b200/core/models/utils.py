"""Step driver — mirrors /root/reference/skyrim/core/models/utils.py:10-49."""
from __future__ import annotations

from datetime import datetime
from typing import Any

import numpy as np
from loguru import logger

from ... import xr_shim as xr


def run_basic_inference(model, n: int, data_source: Any, time: datetime, x=None):
    """Iterate the TimeLoop ``n + 1`` times (the first yield is the initial condition) and stack
    the states into DataArray(time, channel, lat, lon).  The state stays on the GPU between
    steps; each yielded step is copied to the host once (utils.py:37 does ``output.cpu()``)."""
    import torch
    if x is None:
        logger.info("Fetching initial conditions from data source")
        ic = data_source[time]                       # (channel, lat, lon)
        x = torch.from_numpy(np.ascontiguousarray(ic.values, dtype=np.float32))[None, None]
    else:
        logger.info("Using provided initial conditions")
        if isinstance(x, str):
            x = xr.open_dataarray(x)
        vals = np.asarray(x.values, dtype=np.float32)
        if vals.ndim == 3:
            vals = vals[None]
        x = torch.from_numpy(np.ascontiguousarray(vals[-model.n_history_levels:]))[None]  # utils.py:25-31
    arrays, times = [], []
    for k, (t, output, _) in enumerate(model(time, x)):
        arrays.append(output.cpu().numpy().squeeze(0))
        times.append(np.datetime64(t, "s"))
        if k == n:
            break
    stacked = np.stack(arrays)
    coords = dict(time=np.array(times), channel=np.array(model.out_channel_names), lat=np.array(model.grid.lat),
                  lon=np.array(model.grid.lon))
    return xr.DataArray(stacked, dims=["time", "channel", "lat", "lon"], coords=coords)


def estimate_pressure_hpa(elevation_m):
    """Barometric formula (utils.py:52-67)."""
    P0, L, T0, g, M, R = 101325, 0.0065, 288.15, 9.80665, 0.0289644, 8.31447
    return P0 * (1 - (L * elevation_m) / T0) ** (g * M / (R * L)) / 100
