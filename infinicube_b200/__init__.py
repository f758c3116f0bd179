"""infinicube_b200 — B200-native (sm_100a) implementation of InfiniCube's video-generation hot path."""
__version__ = "0.1.0"
