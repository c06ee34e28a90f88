from .threshold_lab import ThresholdLAB  # noqa: F401
__all__ = ['ThresholdLAB']
