"""B200-native Q-learning hot path of uiuc-robovision/video-dqn."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
