from .CCTAContrastCorrector import CCTAContrastCorrector, grid_tiles  # noqa: F401
