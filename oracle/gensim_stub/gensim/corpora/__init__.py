from . import dictionary  # noqa: F401
from .dictionary import Dictionary  # noqa: F401
