"""graphqec_b200 -- B200-native threshold-estimation hot path behind graphqec's Python API.

Package surface mirrors ``/root/reference/graphqec/__init__.py:1-5`` (codes, Measurement,
X_check/Z_check, ThresholdLAB, GF(2) helpers); the Stim/sinter/PyMatching work below
``ThresholdLAB.collect_stats`` runs as sm_100a CUDA kernels behind ``libgqec_b200.so``.
"""

from .circuit import Circuit, target_rec  # noqa: F401
from .dem import DetectorErrorModel  # noqa: F401
from .codes import *  # noqa: F401,F403
from .codes import (  # noqa: F401
    BaseCode,
    RepetitionCode,
    RotatedSurfaceCode,
    UnrotatedSurfaceCode,
    CssCode,
    BivariateBicycleCode,
)
from .measurement import *  # noqa: F401,F403
from .stab import *  # noqa: F401,F403
from .math import *  # noqa: F401,F403
from .stats import Task, TaskStats, AnonTaskStats, wilson_interval  # noqa: F401
from .lab import *  # noqa: F401,F403

__version__ = "0.1.0"
