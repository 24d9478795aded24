"""simudo_b200: B200-native assembly hot path behind Simudo's problem API."""
__version__ = '0.1.0'
