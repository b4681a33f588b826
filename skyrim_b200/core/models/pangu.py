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
        import os
        w = self._weights
        real = os.environ.get("SKYRIM_B200_WEIGHTS")
        if w is None and real:
            # real checkpoint on disk (what pangu.load(registry.get_model("e2mip://pangu")) downloads in the reference,
            # pangu.py:45-46): ONNX initialisers -> engine parameters, no onnx / onnxruntime needed (importers.py)
            from ...importers import load_real_weights
            _, w = load_real_weights("pangu", real)
        eng = StepEngine(self._cfg, self._device)
        # no network here: without SKYRIM_B200_WEIGHTS or a weight dict the engine runs on seeded synthetic weights
        eng.load_weights(w if w is not None else make_pangu_weights(self._cfg, self._seed))
        loop = PanguTimeLoop(eng)
        # a real checkpoint feeds un-normalised residual streams to fp16 tensor-core operands: the first step of every
        # rollout runs with the engine's fp16-range guard (StepEngine.step_guarded raises above 3e4)
        loop.guard_first_step = bool(real) and self._weights is None
        return loop

    @property
    def device(self):
        return self.model.device
