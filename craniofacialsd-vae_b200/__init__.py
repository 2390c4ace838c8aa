"""sdvae_b200 -- B200-native SD-VAE mesh encoder/decoder hot path.

Drop-in for the reference's ``model.py`` (``SpiralConv``, ``Pool``,
``SpiralEnblock``, ``SpiralDeblock``, ``Model``, ``MLPClassifier``) backed by
hand-written sm_100a CUDA kernels behind a C ABI (``include/sdvae_b200.h``).
Sub-modules are imported lazily so that table/fixture code works without CUDA.
"""
__version__ = "0.1.0"

__all__ = ["fixtures", "tables", "cabi", "functional", "model", "losses", "engine"]
