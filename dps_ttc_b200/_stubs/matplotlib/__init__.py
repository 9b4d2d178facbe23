"""Stub: the reference imports matplotlib.pyplot at module scope (gaussian_diffusion.py:4) but the hot path never plots."""
