from .distGCN import DistGCN  # noqa: F401
from .distSAGE import DistSAGE  # noqa: F401
