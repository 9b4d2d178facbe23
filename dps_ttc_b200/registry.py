"""Name → class registries with the reference's API (decorator + getter, NameError on clashes):
  measurements.py:18-32, :225-240     register_operator/get_operator, register_noise/get_noise
  condition_methods.py:8-21           register_conditioning_method/get_conditioning_method
  gaussian_diffusion.py:17-31         register_sampler/get_sampler
"""
from __future__ import annotations


class Registry:
    def __init__(self, what: str):
        self.what = what
        self.table = {}

    def register(self, name: str):
        def wrap(cls):
            if self.table.get(name) is not None:
                raise NameError(f"Name {name} is already registered!")
            self.table[name] = cls
            return cls
        return wrap

    def get(self, name: str):
        if self.table.get(name) is None:
            raise NameError(f"Name {name} is not defined.")
        return self.table[name]

    def names(self):
        return sorted(self.table)


OPERATORS = Registry("operator")
NOISES = Registry("noise")
CONDITIONING = Registry("conditioning method")
SAMPLERS = Registry("sampler")

register_operator = OPERATORS.register
register_noise = NOISES.register
register_conditioning_method = CONDITIONING.register
register_sampler = SAMPLERS.register


def _populate():
    """The classes register themselves on import; make the getters work whatever was imported first."""
    from . import conditioning, operators, sampler  # noqa: F401


def get_operator(name: str, **kwargs):
    _populate()
    return OPERATORS.get(name)(**kwargs)


def get_noise(name: str, **kwargs):
    _populate()
    noiser = NOISES.get(name)(**kwargs)
    noiser.__name__ = name  # read by the conditioning methods (condition_methods.py:35, :50)
    return noiser


def get_conditioning_method(name: str, operator, noiser, **kwargs):
    _populate()
    return CONDITIONING.get(name)(operator=operator, noiser=noiser, **kwargs)


def get_sampler(name: str):
    _populate()
    return SAMPLERS.get(name)
