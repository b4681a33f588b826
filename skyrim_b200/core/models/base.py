"""GlobalModel / GlobalPrediction — mirrors /root/reference/skyrim/core/models/base.py
(adjust_lead_time :13-15, GlobalModel :18-146, GlobalPrediction :149-274,
GlobalPredictionRollout :277-303), with the back-end replaced by the CUDA step engine."""
from __future__ import annotations

import datetime
import time
from pathlib import Path
from typing import List

import numpy as np
from loguru import logger

from ... import xr_shim as xr
from ...common import generate_forecast_id, save_forecast
from .utils import run_basic_inference


def adjust_lead_time(lead_time: int, step_size: int = 6):
    """Adjust lead time to the nearest multiple of step_size (base.py:13-15)."""
    return max(step_size, (lead_time // step_size) * step_size)


class SyntheticDataSource:
    """Seeded synthetic initial conditions.  The reference's sources
    (cds / gfs / ifs, libs/ic/__init__.py:25-34) download GRIB over the network and stay with
    the reference package; any object with ``source[time] -> DataArray(channel, lat, lon)`` works."""

    def __init__(self, channel_names, nlat=721, nlon=1440, seed=0):
        self.channel_names, self.nlat, self.nlon, self.seed = list(channel_names), nlat, nlon, seed

    def __getitem__(self, t):
        from ...weights import synthetic_state
        from ...timeloop import equiangular_grid
        g = equiangular_grid(self.nlat, self.nlon)
        x = synthetic_state(self.channel_names, self.nlat, self.nlon, self.seed)
        return xr.DataArray(x, dims=["channel", "lat", "lon"],
                            coords=dict(channel=np.array(self.channel_names), lat=np.array(g.lat), lon=np.array(g.lon)))


class GlobalModel:
    def __init__(self, model_name: str, ic_source: str = "synthetic", **engine_kw):
        clock = time.time()
        self.model_name = model_name
        self.ic_source = ic_source
        self.engine_kw = engine_kw
        logger.debug(f"Building {model_name} model...")
        self.model = self.build_model()
        logger.debug(f"Building {self.ic_source} data source...")
        self.data_source = self.build_datasource()
        logger.success(f"Initialized {model_name} in {time.time() - clock:.1f} seconds")

    def build_model(self):
        raise NotImplementedError

    def build_datasource(self):
        if self.ic_source == "synthetic":
            return SyntheticDataSource(self.model.in_channel_names, len(self.model.grid.lat), len(self.model.grid.lon))
        if hasattr(self.ic_source, "__getitem__") and not isinstance(self.ic_source, str):
            return self.ic_source  # a user-supplied data source object
        raise NotImplementedError(
            f"ic_source='{self.ic_source}' downloads initial conditions over the network (reference "
            "libs/ic/__init__.py:25-34); pass ic_source='synthetic' or a data-source object")

    def release_model(self):
        self.model.engine.close()

    @property
    def time_step(self):
        return self.model.time_step

    def time_steps(self, lead_time: int):
        lead_time = adjust_lead_time(lead_time, step_size=6)
        return int(lead_time // (self.time_step.total_seconds() / 3600))

    @property
    def in_channel_names(self):
        return self.model.in_channel_names

    @property
    def out_channel_names(self):
        return self.model.out_channel_names

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(model_name={self.model_name})"

    def predict_one_step(self, start_time: datetime.datetime, initial_condition=None):
        return run_basic_inference(model=self.model, n=1, data_source=self.data_source, time=start_time,
                                   x=initial_condition)

    def forecast(self, start_time: datetime.datetime, n_steps: int = 3, channels: List[str] = []):
        """All steps from the IC in one device-resident generator (base.py:94-117)."""
        da = run_basic_inference(model=self.model, n=n_steps, data_source=self.data_source, time=start_time, x=None)
        return da.sel(channel=list(channels)) if channels else da

    def rollout(self, start_time: datetime.datetime, n_steps: int = 3, save: bool = True, save_config: dict = {},
                initial_condition=None):
        """Chained 6-h steps; returns (last prediction with 2 time slices, paths) like base.py:119-146.
        Unlike the reference (which round-trips the state through the host and re-uploads it every
        step, utils.py:24-31), the state stays resident in HBM; only what is saved / returned is
        copied back.  ``initial_condition`` (DataArray / path) resumes a rollout (TODO at base.py:127)."""
        import torch
        pred, output_paths, source = None, [], self.ic_source if isinstance(self.ic_source, str) else "file"
        save_config = dict(save_config)
        save_config.setdefault("forecast_id", generate_forecast_id())
        if initial_condition is None:
            ic = self.data_source[start_time]
        else:
            ic = xr.open_dataarray(initial_condition) if isinstance(initial_condition, str) else initial_condition
            source = "file"
        vals = np.asarray(ic.values, dtype=np.float32)
        vals = vals[-1] if vals.ndim == 4 else vals
        x = torch.from_numpy(np.ascontiguousarray(vals))[None, None]
        if save and hasattr(self.model, "iter_host"):
            # every state is needed on the host: device->host copy of step n overlaps step n+1 (timeloop.iter_host)
            gen = ((t, h[0].numpy(), None) for t, h in self.model.iter_host(start_time, x, n_steps))
            to_host = lambda a: a
        else:
            gen = self.model(start_time, x)
            to_host = lambda a: a.cpu().numpy()[0]
        t_prev, prev, _ = next(gen)
        prev_host = np.array(to_host(prev))   # own copy: the iterator's host buffers are a ring of two
        for n in range(n_steps):
            t_cur, cur, _ = next(gen)
            need_host = save or n == n_steps - 1
            if need_host:
                cur_host = to_host(cur)
                pred = xr.DataArray(np.stack([prev_host, cur_host]), dims=["time", "channel", "lat", "lon"],
                                    coords=dict(time=np.array([np.datetime64(t_prev, "s"), np.datetime64(t_cur, "s")]),
                                                channel=np.array(self.out_channel_names),
                                                lat=np.array(self.model.grid.lat), lon=np.array(self.model.grid.lon)))
                if save:
                    output_paths.append(save_forecast(pred, self.model_name, t_prev, t_cur, source, config=save_config))
                prev_host = pred.values[1]
            t_prev, source = t_cur, "file"
            logger.success(f"Rollout step {n+1}/{n_steps} completed")
        return pred, output_paths


class GlobalPrediction:
    filepath: Path | None = None
    prediction = None

    def __init__(self, source, model_name: str = ""):
        self.model = model_name
        if isinstance(source, (str, Path)):
            self.filepath = Path(source)
            self.prediction = xr.open_dataarray(source).squeeze()
        elif hasattr(source, "dims") and hasattr(source, "values"):
            self.filepath = None
            self.prediction = source.squeeze()
        else:
            raise ValueError("Invalid source type.")

    @property
    def coords(self):
        return self.prediction.coords

    @property
    def size(self):
        return self.prediction.size

    @property
    def channels(self):
        return self.prediction.channel

    def __repr__(self) -> str:
        info = self.filepath if self.filepath else f"{type(self.prediction).__name__} with shape {self.prediction.shape}"
        return f"GlobalPrediction(model={self.model},source={info})"

    def slice(self, lat: slice | None = None, lon: slice | None = None, channel: str | None = None,
              n_step: slice | None = None):
        if channel is None:
            data = self.prediction
        else:
            assert channel in self.channels, f"Variable {channel} not found in dataset."
            data = self.prediction.sel(channel=channel)
        if lat:
            data = data.sel(lat=lat)
        if lon:
            data = data.sel(lon=lon)
        if n_step and "time" in data.dims:
            data = data.isel(time=n_step)
        return data

    def point(self, lat: float, lon: float, channel: str, n_step: int | None = 1):
        if lon < 0:
            lon = 360 + lon
        assert channel in self.channels, f"Variable {channel} not found in dataset."
        lats, lons = np.asarray(self.prediction.coords["lat"]), np.asarray(self.prediction.coords["lon"])
        if lat not in lats or lon not in lons:
            lat = float(lats[np.abs(lats - lat).argmin()])
            lon = float(lons[np.abs(lons - lon).argmin()])
            logger.warning(f"Exact coordinates not found. Using nearest values: Lat {lat}, Lon {lon}")
        data = self.prediction.sel(lat=lat, lon=lon, channel=channel)
        if "time" in data.dims:
            data = data.isel(time=n_step)
        return data.item()

    def point_wind_uv(self, lat: float, lon: float, pressure_level: int = 1000, n_step: int | None = 1):
        u = self.point(lat=lat, lon=lon, channel=f"u{pressure_level}", n_step=n_step)
        v = self.point(lat=lat, lon=lon, channel=f"v{pressure_level}", n_step=n_step)
        return u, v

    def wind_speed(self, lat: float, lon: float, pressure_level: int, n_step: int | None = 1):
        u, v = self.point_wind_uv(lat, lon, pressure_level, n_step)
        return (u ** 2 + v ** 2) ** 0.5

    def surface_wind_speed(self, lat: float, lon: float, n_step: int | None = 1):
        return self.wind_speed(lat, lon, pressure_level=1000, n_step=n_step)


class GlobalPredictionRollout:
    def __init__(self, rollout: list):
        self.rollout = [GlobalPrediction(source) for source in rollout]
        self.time_steps = [np.asarray(r.prediction.coords["time"])[-1] for r in self.rollout]

    def __repr__(self):
        return f"<GlobalPredictionRollout with {len(self.rollout)} predictions, last times: {self.time_steps}>"

    def wind_speed(self, lat: float, lon: float, pressure_level: int, n_step: int | None = 1):
        return [pred.wind_speed(lat, lon, pressure_level, n_step) for pred in self.rollout]

    def surface_wind_speed(self, lat: float, lon: float, n_step: int | None = 1):
        return self.wind_speed(lat, lon, pressure_level=1000, n_step=n_step)
