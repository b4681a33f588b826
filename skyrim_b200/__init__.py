"""skyrim_b200 — B200-native rollout engine behind the Skyrim API."""
__all__ = ["Skyrim"]


def __getattr__(name):
    if name == "Skyrim":
        from .core.skyrim import Skyrim
        return Skyrim
    raise AttributeError(name)
