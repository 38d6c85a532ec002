"""HU scalers with the reference's class names (reference data/Scaler.py:20-45)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class ZeroCenterScaler:
    low: int
    high: int
    shift: int = field(init=False, default=None)

    def __post_init__(self):
        self.shift = (self.high - abs(self.low)) // 2  # 238 for (-1024, 1500)

    def __call__(self, x):
        return x - self.shift

    def unscale(self, x):
        return x + self.shift


@dataclass
class FactorZeroCenterScaler(ZeroCenterScaler):
    factor: int = 1

    def __call__(self, x):
        return super().__call__(x) / self.factor

    def unscale(self, x):
        return super().unscale(x * self.factor)
