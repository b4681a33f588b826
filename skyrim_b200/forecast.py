"""`forecast` command line — mirrors /root/reference/skyrim/forecast.py:19-151 (same option names, short flags and
defaults; same `run_forecast` signature and return value) on top of the B200 engine:

    python -m skyrim_b200.forecast -m pangu -d 20240507 -t 0000 -l 24 -o outputs

Differences, all forced by the scope of this build: `--initial_conditions` accepts `synthetic`
(the default here: the network IC providers cds / ifs / gfs are the reference's `libs/ic`, out of scope, and raise a clear
error) or a path to a netCDF / .npy state; `--modal` (remote execution through Modal) is not available and says so;
`--weight_seed`, `--device` and `--grid` select the synthetic weight seed, the GPU and a coarse test grid."""
from __future__ import annotations

from datetime import datetime, timedelta
from pathlib import Path

import click
from loguru import logger

from .common import AVAILABLE_MODELS

yesterday = (datetime.now() - timedelta(days=1)).date().isoformat().replace("-", "")


def run_forecast(model_name: str, date: str, time: str, lead_time: int, list_models: bool, initial_conditions: str,
                 output_dir: str, filter_vars: str, **model_kw):
    from .core import Skyrim
    if list_models:
        print("Available models:", Skyrim.list_available_models())
        return []
    logger.debug(f"Starting forecast run with model_name={model_name}, date={date}, time={time}, lead_time={lead_time}, "
                 f"initial_conditions={initial_conditions}, output_dir={output_dir}, filter_vars={filter_vars}")
    model = Skyrim(model_name, ic_source=initial_conditions, **model_kw)
    pred, output_paths = model.predict(
        date=date, time=time, lead_time=lead_time, save=True,
        save_config={"output_dir": output_dir or str(Path.cwd() / "outputs"),
                     "filter_vars": [v.strip() for v in filter_vars.split(",") if v.strip()]})
    return output_paths


@click.command()
@click.option("--model_name", "-m", type=click.Choice(AVAILABLE_MODELS, case_sensitive=False), default="pangu", help="Select model")
@click.option("--date", "-d", type=str, default=yesterday, help="YYYYMMDD")
@click.option("--time", "-t", type=str, default="0000", help="HHMM")
@click.option("--lead_time", "-l", type=int, default=6, help="Lead time in hours (floored to a multiple of 6)")
@click.option("--list_models", "-lm", is_flag=True, help="List all available models and exit")
@click.option("--initial_conditions", "-ic", type=str, default="synthetic",
              help="Initial conditions: 'synthetic' (seeded state) or a path to a netCDF / .npy state; cds / ifs / gfs need the "
                   "reference's network data sources")
@click.option("--output_dir", "-o", type=str, default="", help="Output directory (local)")
@click.option("--filter_vars", "-f", type=str, default="", help="Filter variables such as t2m (temperature) before saving forecasts.")
@click.option("--modal", "-mo", is_flag=True, help="(reference option) run on Modal — not available in this build")
@click.option("--weight_seed", type=int, default=0, help="seed of the synthetic weights (no checkpoint download here)")
@click.option("--device", type=int, default=0, help="CUDA device ordinal")
@click.option("--grid", type=str, default="", help="NLATxNLON coarse test grid (default: the model's 721x1440)")
def main(model_name, date, time, lead_time, list_models, initial_conditions, output_dir, filter_vars, modal, weight_seed, device, grid):
    if modal:
        raise click.UsageError("--modal: remote execution through Modal is outside this build (SURVEY.md section 2); run locally")
    kw = {"weight_seed": weight_seed, "device": device}
    if grid:
        from .config import graphcast_small, pangu_small, sfno_small
        nlat, nlon = (int(v) for v in grid.lower().split("x"))
        kw["cfg"] = (pangu_small(nlat, nlon) if model_name == "pangu" else graphcast_small(nlat, nlon) if model_name == "graphcast"
                     else sfno_small(nlat, nlon))
    paths = run_forecast(model_name, date, time, lead_time, list_models, initial_conditions, output_dir, filter_vars, **kw)
    for p in paths:
        print(p)
    return paths


if __name__ == "__main__":
    main()
