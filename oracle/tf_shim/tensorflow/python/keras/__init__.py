from . import regularizers  # noqa: F401
