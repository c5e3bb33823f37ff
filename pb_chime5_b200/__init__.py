"""B200-native guided source separation (drop-in for pb_chime5.core)."""
__version__ = '0.1.0'
