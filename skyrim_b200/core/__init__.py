from .skyrim import Skyrim  # noqa: F401
from .models import MODELS  # noqa: F401
