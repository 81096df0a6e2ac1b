from .model import Model, log  # noqa: F401
