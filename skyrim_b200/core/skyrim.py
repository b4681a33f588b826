"""Facade — mirrors /root/reference/skyrim/core/skyrim.py:12-95 (same constructor, ``predict``,
``forecast``, ``list_available_models``)."""
from __future__ import annotations

import datetime

from loguru import logger

from .models import MODELS
from .models.base import GlobalModel, GlobalPrediction, adjust_lead_time


class Skyrim:
    def __init__(self, *model_names: str, ic_source="synthetic", **model_kw):
        missing_names = [name for name in model_names if name not in MODELS]
        if missing_names:
            raise ValueError(f"Invalid model name(s): {missing_names}")
        if len(model_names) != 1:
            # the reference's multi-model mean (ensemble.py:10-133) is broken at HEAD and out of scope;
            # perturbed-IC ensembles of ONE model live in skyrim_b200.ensemble
            raise NotImplementedError("exactly one model name is supported; see skyrim_b200.ensemble")
        self.model_names = model_names
        self.ic_source = ic_source
        self.model: GlobalModel = MODELS[model_names[0]](ic_source=ic_source, **model_kw)
        logger.debug(f"Initialized {self.model} model with initial conditions from {ic_source}")

    def __repr__(self) -> str:
        return f"Skyrim(models={self.model_names},ic={self.ic_source})"

    @staticmethod
    def list_available_models():
        return list(MODELS.keys())

    def forecast(self, start_time: datetime.datetime, n_steps: int = 4, channels: list = []):
        start_time = start_time.replace(second=0, microsecond=0)
        return self.model.forecast(start_time=start_time, n_steps=n_steps, channels=channels)

    def predict(self, date: str, time: str, lead_time: int = 6, save: bool = False, save_config: dict = {}):
        """Predict a single lead-time snapshot, optionally saving every intermediate step
        (skyrim.py:60-95: YYYYMMDD / HHMM parsing, lead time floored to a multiple of 6 h)."""
        start_time = datetime.datetime(int(date[:4]), int(date[4:6]), int(date[6:8]), int(time[:2]), int(time[2:4]))
        lead_time = adjust_lead_time(lead_time, step_size=6)
        logger.debug(f"Lead time adjusted to nearest multiple of 6: {lead_time} hours")
        n_steps = int(lead_time // (self.model.time_step.total_seconds() / 3600))
        logger.debug(f"Number of prediction steps: {n_steps}")
        pred, output_paths = self.model.rollout(start_time=start_time, n_steps=n_steps, save=save,
                                                save_config=save_config)
        logger.debug("Prediction completed successfully")
        return GlobalPrediction(pred, model_name=self.model_names), output_paths
