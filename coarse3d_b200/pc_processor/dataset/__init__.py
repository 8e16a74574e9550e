from . import preprocess  # noqa: F401
