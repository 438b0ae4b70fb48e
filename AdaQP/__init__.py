"""Drop-in alias: `from AdaQP import Trainer` (the reference's main.py:3) resolves to the
B200-native package.  Sub-packages are aliased too (AdaQP.communicator, AdaQP.manager, ...)."""
import importlib
import sys

import adaqp_b200 as _impl

for _name in ("helper", "util", "communicator", "manager", "assigner", "model", "trainer"):
    sys.modules[f"AdaQP.{_name}"] = importlib.import_module(f"adaqp_b200.{_name}")
from adaqp_b200.trainer import Trainer  # noqa: E402,F401
