from .contrast_pixel_loss import ContrastMEMLoss  # noqa: F401
from .lovasz_softmax import Lovasz_softmax, lovasz_softmax  # noqa: F401
