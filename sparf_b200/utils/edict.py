"""`edict`: the attribute-dict type the reference's API traffics in (easydict.EasyDict).
Uses the real package when installed, else a small equivalent (same semantics the reference relies
on: attribute + item access, nested wrapping, hasattr False for missing keys)."""
try:  # pragma: no cover
    from easydict import EasyDict as edict
except ImportError:

    class edict(dict):
        def __init__(self, d=None, **kwargs):
            super().__init__()
            d = dict(d or {}, **kwargs)
            for k, v in d.items():
                setattr(self, k, v)

        @classmethod
        def _wrap(cls, v):
            if isinstance(v, dict) and not isinstance(v, edict):
                return cls(v)
            if isinstance(v, (list, tuple)):
                return type(v)(cls._wrap(x) for x in v)
            return v

        def __setattr__(self, k, v):
            v = self._wrap(v)
            super().__setattr__(k, v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def update(self, e=None, **f):
            for k, v in dict(e or {}, **f).items():
                setattr(self, k, v)

        def pop(self, k, *a):
            if hasattr(self, k):
                super().__delattr__(k)
            return super().pop(k, *a)
