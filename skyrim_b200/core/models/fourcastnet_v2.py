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
        import os
        real = os.environ.get("SKYRIM_B200_WEIGHTS_FCNV2")
        guard = self._weights is None and bool(real)
        if self._weights is None and real:
            # fcnv2_sm checkpoint directory (weights.tar + global_means.npy / global_stds.npy: what fcnv2_sm.load reads in
            # the reference, fourcastnet_v2.py:36-37); hyper-parameters come from the tensor shapes
            from ...importers import load_real_weights
            self._cfg, self._weights = load_real_weights("sfno", real)
        eng = StepEngine(self._cfg, self._device)
        w = dict(self._weights if self._weights is not None else make_sfno_weights(self._cfg, self._seed))
        w.update(sfno_tables(self._cfg))
        eng.load_weights(w)
        loop = SFNOTimeLoop(eng)
        loop.guard_first_step = guard   # real checkpoint: first step of every rollout under the fp16-range guard
        return loop
