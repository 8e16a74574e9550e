from .prototype import PrototypeBank, l2_normalize, momentum_update, prototype_learning  # noqa: F401
