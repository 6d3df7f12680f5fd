"""Minimal stand-in for `gymnasium`, used ONLY by tests/golden/make_golden.py so that the
reference package (read-only at /root/reference) can be imported in the build container.
Never imported by the product or by any test."""
from . import spaces, vector, envs  # noqa: F401


class Env:  # pragma: no cover - annotation target only
    pass


class Wrapper(Env):  # pragma: no cover
    pass


class VectorizeMode:  # pragma: no cover
    SYNC = "sync"
    ASYNC = "async"
    VECTOR_ENTRY_POINT = "vector_entry_point"


def make(*args, **kwargs):  # pragma: no cover
    raise RuntimeError("gymnasium shim: no environments available")


make_vec = make
