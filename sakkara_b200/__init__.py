"""sakkara_b200: B200-native evaluator of logp + dlogp for sakkara-built hierarchical models."""
__version__ = '0.1.0'
