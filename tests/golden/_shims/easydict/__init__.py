"""Minimal stand-in for the `easydict` package (not installed in this image).

Only used (a) by tests/golden/make_golden.py to import the UNMODIFIED reference
from /root/reference in the build container, and (b) as the dict type of our own
host-side mirror when the real package is absent.  Behaviour needed by the
reference: attribute + item access, recursive wrapping of nested dicts,
hasattr() False for missing keys, update()/pop() keeping both views in sync.
"""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        if d is None:
            d = {}
        if kwargs:
            d = dict(d, **kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _wrap(cls, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            return cls(value)
        if isinstance(value, (list, tuple)):
            return type(value)(cls._wrap(x) for x in value)
        return value

    def __setattr__(self, name, value):
        value = self._wrap(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = e or dict()
        d = dict(d, **f) if f else d
        for k in d:
            setattr(self, k, d[k])

    def pop(self, k, *args):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *args)

    def __delattr__(self, name):
        super().__delattr__(name)
        if name in self:
            super().__delitem__(name)
