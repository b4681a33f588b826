"""FourcastnetV2Model — mirrors /root/reference/skyrim/core/models/fourcastnet_v2.py:22-49."""
from __future__ import annotations

from ...config import FCNV2_CHANNELS, SFNOConfig, sfno_full
from .base import GlobalModel

CHANNELS = FCNV2_CHANNELS


class FourcastnetV2Model(GlobalModel):
    model_name = "fourcastnet_v2"

    def __init__(self, *args, cfg: SFNOConfig | None = None, weights=None, weight_seed: int = 0, device: int = 0,
                 **kwargs):
        self._cfg, self._weights, self._seed, self._device = cfg or sfno_full(), weights, weight_seed, device
        super().__init__(self.model_name, *args, **kwargs)

    def build_model(self):
        from ...engine import StepEngine
        from ...timeloop import SFNOTimeLoop
        from ...weights import make_sfno_weights, sfno_tables
        eng = StepEngine(self._cfg, self._device)
        w = dict(self._weights if self._weights is not None else make_sfno_weights(self._cfg, self._seed))
        w.update(sfno_tables(self._cfg))
        eng.load_weights(w)
        return SFNOTimeLoop(eng)
