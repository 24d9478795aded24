"""``SetattrInitMixin`` -- same contract as ``simudo/util/setattr_init_mixin.py:4-8``."""

__all__ = ['SetattrInitMixin']


class SetattrInitMixin(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
