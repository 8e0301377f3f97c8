"""Minimal stand-in for the `easydict` package (absent offline), used ONLY by
tests/golden/make_golden.py to import the reference on CPU.  Reproduces the
behaviour the reference relies on: attribute access, dict -> EasyDict and
list/tuple -> list conversion on attribute set (lib/vnlb/proc_nl.py:175)."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = [self.__class__(x) if isinstance(x, dict) else x for x in value]
        elif isinstance(value, dict) and not isinstance(value, self.__class__):
            value = self.__class__(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = e or dict()
        d.update(f)
        for k in d:
            setattr(self, k, d[k])

    def pop(self, k, d=None):
        delattr(self, k)
        return super().pop(k, d)
