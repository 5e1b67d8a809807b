from . import sampling  # noqa: F401
