"""PanguModel — mirrors /root/reference/skyrim/core/models/pangu.py:17-62; ``build_model``
returns the CUDA TimeLoop instead of ``pangu.load(registry.get_model("e2mip://pangu"))``."""
from __future__ import annotations

from ...config import PANGU_CHANNELS, PanguConfig, pangu_full
from .base import GlobalModel

CHANNELS = PANGU_CHANNELS


class PanguModel(GlobalModel):
    model_name = "pangu"

    def __init__(self, *args, cfg: PanguConfig | None = None, weights=None, weight_seed: int = 0, device: int = 0,
                 **kwargs):
        self._cfg, self._weights, self._seed, self._device = cfg or pangu_full(), weights, weight_seed, device
        super().__init__(self.model_name, *args, **kwargs)

    def build_model(self):
        from ...engine import StepEngine
        from ...timeloop import PanguTimeLoop
        from ...weights import make_pangu_weights
        eng = StepEngine(self._cfg, self._device)
        # no network: real checkpoints cannot be downloaded here, the engine runs on seeded
        # synthetic weights unless a weight dict is supplied
        eng.load_weights(self._weights if self._weights is not None else make_pangu_weights(self._cfg, self._seed))
        return PanguTimeLoop(eng)

    @property
    def device(self):
        return self.model.device
