from . import path, simulator  # noqa: F401
