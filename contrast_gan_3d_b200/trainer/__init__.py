from .Trainer import Trainer  # noqa: F401
