"""dohp-b200: B200-native implementation of Dohp's matrix-free high-order element hot path."""
__version__ = "0.1.0"
