from .generate_correlation import generate_correlation

__all__ = ['generate_correlation']
