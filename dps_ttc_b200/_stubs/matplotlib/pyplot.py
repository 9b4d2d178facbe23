"""Stub of matplotlib.pyplot: every call is a no-op."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None
    return _noop
