"""Forecast output: mirrors /root/reference/skyrim/common.py for LOCAL targets
(generate_forecast_id :23-31, SaveConfig :34-45, generate_filename :48-69, save_forecast
:115-204).  Remote targets (s3://, hf://) need network credentials and the reference's own
boto3 / huggingface paths; they raise here (out of scope).

Two reference quirks are fixed consciously and documented in DESIGN.md:
  * common.py:126-129 silently turns every request without an explicit file_type into zarr
    because ``target`` is always truthy — here a local target defaults to netCDF, as the
    reference's own README/notebook output (28 ``.nc`` files) shows was intended;
  * common.py:150 appends local zarr along "step" although the data's dim is "time" (fails at
    HEAD) — here the append dimension is "time".
"""
from __future__ import annotations

import hashlib
import os
import time
from dataclasses import dataclass, field
from datetime import datetime
from pathlib import Path
from typing import Callable
from urllib.parse import urlparse

from loguru import logger

AVAILABLE_MODELS = ["pangu", "fourcastnet_v2", "graphcast"]
LOCAL_CACHE = os.path.join(os.path.expanduser("~"), ".cache", "skyrim")
OUTPUT_DIR = str(Path.cwd() / "outputs")

_B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def _b58encode(b: bytes) -> str:
    n = int.from_bytes(b, "big")
    out = ""
    while n:
        n, r = divmod(n, 58)
        out = _B58[r] + out
    return "1" * (len(b) - len(b.lstrip(b"\0"))) + out


def generate_forecast_id(length=10):
    """sha256(time) -> base58[:length]  (common.py:23-31; base58 is not installed here)."""
    return _b58encode(hashlib.sha256(str(time.time()).encode()).digest())[:length]


@dataclass
class SaveConfig:
    forecast_id: str = ""
    output_dir: str = OUTPUT_DIR
    file_type: str = "netcdf"
    filter_vars: tuple = ()
    mapping_func: Callable = lambda x: x
    zarr_store_config: dict = field(default_factory=dict)

    def __post_init__(self):
        if not self.forecast_id:
            self.forecast_id = generate_forecast_id()


def generate_filename(model: str, start_time: datetime, pred_time: datetime, ic_source: str = "cds"):
    # common.py:62-69
    return (f"{model}__{ic_source}__{start_time.strftime('%Y%m%d_%H:%M')}__"
            f"{pred_time.strftime('%Y%m%d_%H:%M')}.nc")


def save_forecast(pred, model_name: str, start_time: datetime, pred_time: datetime, source: str = "cds",
                  config: dict = {}):
    config = SaveConfig(**config)
    p = urlparse(config.output_dir)
    target = p.scheme or "local"
    if target != "local":
        raise NotImplementedError(f"remote output target '{target}' needs network access; only local paths are built")
    pred = config.mapping_func(pred)
    pred = pred[list(config.filter_vars)] if len(config.filter_vars) else pred
    if config.file_type == "netcdf":
        filename = generate_filename(model_name, start_time, pred_time, source)
        output_path = Path(config.output_dir) / config.forecast_id / filename
        logger.info(f"Saving outputs to {output_path}")
        output_path.parent.mkdir(parents=True, exist_ok=True)
        pred.to_netcdf(output_path, engine="scipy")
    elif config.file_type == "zarr":
        output_path = str(Path(config.output_dir) / config.forecast_id)
        if Path(output_path).exists():
            pred.to_zarr(output_path, append_dim="time", mode="a", consolidated=True)
        else:
            pred.to_zarr(output_path, mode="w", consolidated=True)
    else:
        raise ValueError(f"Invalid file type. {config.file_type} not supported.")
    logger.success(f"Results saved to: {output_path}")
    return str(output_path)
