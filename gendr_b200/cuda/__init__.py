from . import generalized_renderer  # noqa: F401
