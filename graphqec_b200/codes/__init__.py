from .base_code import BaseCode  # noqa: F401
from .repetition_code import RepetitionCode  # noqa: F401
from .rotated_surface_code import RotatedSurfaceCode  # noqa: F401
from .unrotated_surface_code import UnrotatedSurfaceCode  # noqa: F401
from .css_code import CssCode, compute_logicals  # noqa: F401
from .bivariate_bicycle_code import BivariateBicycleCode  # noqa: F401
