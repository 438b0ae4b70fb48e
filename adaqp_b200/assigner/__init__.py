from .assigner import Assigner  # noqa: F401
