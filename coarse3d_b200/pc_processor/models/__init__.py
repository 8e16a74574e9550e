from .prototype import PrototypeBank, l2_normalize, momentum_update  # noqa: F401
