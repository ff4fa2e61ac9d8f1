"""B200-native CLIP dual-encoder hot path (hand-written sm_100a kernels behind the `clip` API)."""
__version__ = "0.1.0"
