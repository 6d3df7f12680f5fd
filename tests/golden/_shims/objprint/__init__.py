"""Minimal stand-in for `objprint` (see gymnasium shim docstring)."""


def add_objprint(cls=None, **kwargs):
    if cls is None:
        return lambda c: c
    return cls


def objstr(obj, **kwargs):
    return repr(obj)


def op(*args, **kwargs):
    print(*args)
