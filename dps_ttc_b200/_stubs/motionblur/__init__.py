"""Stub of the un-vendored `motionblur` package (LeviBorodenko/motionblur, unpinned in the reference's Dockerfile:25)."""
