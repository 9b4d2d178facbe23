"""dps_ttc_b200 — B200-native kernels behind the dps-ttc plugin surface."""
__version__ = "0.1.0"
