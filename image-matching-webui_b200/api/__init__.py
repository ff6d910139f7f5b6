from .core import ImageMatchingAPI  # noqa: F401
