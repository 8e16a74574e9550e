from .contrast_pixel_loss import ContrastMEMLoss  # noqa: F401
