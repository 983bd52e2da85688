from .samplers import BaseSampler, PseudoSampler, RandomSampler, SamplingResult

__all__ = ['BaseSampler', 'PseudoSampler', 'RandomSampler', 'SamplingResult']
