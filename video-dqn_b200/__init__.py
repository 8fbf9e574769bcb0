"""B200-native Q-learning hot path of uiuc-robovision/video-dqn (package directory
`video-dqn_b200`; import it as `video_dqn_b200`)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
