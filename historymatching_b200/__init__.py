"""B200-native history matching: ensemble TPFA forward runs and ensemble-smoother
analysis (ES / ES-MDA / IES / localised) behind the notebooks' own Python surface.

Call ``historymatching_b200.activate()`` before the notebook's imports to make
``import TPFA_ResSim`` and ``from tools import geostat, utils`` resolve to the
B200-native drop-ins in ``historymatching_b200/dropin``.
"""

from __future__ import annotations

import os
import sys

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def activate():
    """Put the drop-in ``TPFA_ResSim`` and ``tools`` packages first on ``sys.path``."""
    if DROPIN_DIR not in sys.path:
        sys.path.insert(0, DROPIN_DIR)
    for name in ("TPFA_ResSim", "tools"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(DROPIN_DIR):
            raise ImportError(f"a different '{name}' module is already imported from {mod.__file__}")
    return DROPIN_DIR
