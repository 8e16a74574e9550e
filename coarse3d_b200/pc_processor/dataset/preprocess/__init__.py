from . import projection  # noqa: F401
from .projection import RangeProjection  # noqa: F401
