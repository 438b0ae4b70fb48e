from .recorder import Recorder  # noqa: F401
from .timer import Timer  # noqa: F401
