"""Import-only stub (matplotlib is absent offline; never called on the hot path)."""


def use(*a, **k):
    pass
