from .likelihood import Likelihood
from ._direct_likelihood import DirectLikelihood
from ._profile_likelihood import ProfileLikelihood

__all__ = ['Likelihood', 'DirectLikelihood', 'ProfileLikelihood']
