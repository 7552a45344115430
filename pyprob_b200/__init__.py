"""pyprob_b200 — B200-native inference-compilation hot path of pyprob."""
__version__ = '0.1.0'
