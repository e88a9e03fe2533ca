"""B200-native (sm_100a) implementation of the scattered-to-grid PTV interpolation hot path of
tombultreys/ptv_interpolation, behind the reference's own Python API.

    from ptv_interpolation_b200 import interpolator, physics      # drop-in modules
    from ptv_interpolation_b200.engine import PTVEngine             # device-resident engine

The compute path is hand-written CUDA in ``csrc/`` reached through the C ABI of
``include/ptv_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
