"""Model registry (reference: skyrim/core/models/__init__.py:9-17).  The B200 engine builds the
two step operators the north star names; the other reference wrappers (fourcastnet, dlwp,
graphcast, fuxi, fengwu) are not part of this path."""
from .base import GlobalModel, GlobalPrediction, GlobalPredictionRollout  # noqa: F401
from .fourcastnet_v2 import FourcastnetV2Model
from .pangu import PanguModel

MODELS = {"pangu": PanguModel, "fourcastnet_v2": FourcastnetV2Model}
