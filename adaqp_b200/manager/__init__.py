from .graphEngine import DecompGraph, GraphEngine  # noqa: F401
