"""jaxili_b200: B200-native implementation of jaxili's ConditionalMAF hot path."""

__version__ = "0.1.0"
