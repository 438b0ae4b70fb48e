from .typing import DistGNNType, BitType, MessageType, ProprogationMode  # noqa: F401
