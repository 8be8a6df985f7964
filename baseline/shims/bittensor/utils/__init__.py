from . import networking  # noqa: F401
