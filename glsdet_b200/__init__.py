"""glsdet_b200: B200-native (sm_100a) GLSDet neck + head + decode + NMS behind the reference's module API."""
__version__ = "0.1.0"
