from .mixed_correlation import MixedCorrelation

__all__ = ['MixedCorrelation']
